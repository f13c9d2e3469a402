"""Refresh profiles/sc_counters.json (GPU box): warp instructions per codeword and DRAM bytes per launch of the two
kernels bench.py quotes a roofline for, captured with ncu from the CURRENT sources and stamped with their hash
(bench.py refuses a capture whose stamp differs from the sources it runs).
   python tools/refresh_counters.py            -> gpurun_out/sc_counters.json   (copy to profiles/ after review)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import source_stamp  # noqa: E402

METRICS = "smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"


def capture(key, what, match, batch):
    cmd = ["ncu", "--metrics", METRICS, "--clock-control", "none", "--csv", "-k", "regex:" + match, "-c", "2",
           sys.executable, os.path.join(ROOT, "tools", "ncu_target.py"), what, "2"]
    raw = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(raw)) if len(r) > 10]
    hdr = rows[0]
    iid, im, iv, ik = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Kernel Name")
    last = {}
    for r in rows[1:]:
        last.setdefault(r[iid], {"kernel": r[ik]})[r[im]] = float(r[iv].replace(",", ""))
    m = last[sorted(last, key=int)[-1]]                    # second (warm) launch
    return {"source_stamp": source_stamp(key), "kernel": m["kernel"], "batch": batch, "warp_instr_per_codeword": m["smsp__inst_executed.sum"] / batch,
            "dram_bytes_per_launch": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"],
            "dram_bytes_read": m["dram__bytes_read.sum"], "dram_bytes_write": m["dram__bytes_write.sum"],
            "ncu_time_ms": m["gpu__time_duration.sum"] / 1e6}


def main():
    out = {"how": "ncu --metrics %s --clock-control none (tools/refresh_counters.py), second launch" % METRICS,
           "sc5_kernel<10>": capture("sc5_kernel<10>", "sc", "sc5_kernel", 1 << 20),
           "scl3_kernel<10,8>": capture("scl3_kernel<10,8>", "scl", "scl3_kernel", 1 << 18)}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sc_counters.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
