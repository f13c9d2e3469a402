"""Summarise an .ncu-rep (run here, no GPU needed) into a small markdown file under profiles/.
   python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r01_x.md "title" """
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum.per_cycle_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["# " + title, "", "source: `%s` (ncu --set full --clock-control none --import-source on)" % rep, ""]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines += ["## kernel `%s`" % d.get("Kernel Name", "?"), "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in d:
                lines.append("| %s | %s | %s |" % (k, d[k], units[hdr.index(k)]))
        lines += ["", "stall reasons (warps per issue-active cycle):", "", "| reason | value |", "|---|---|"]
        st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(d[h] or 0))
              for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
        for nm, v in sorted(st, key=lambda t: -t[1])[:10]:
            lines.append("| %s | %.3f |" % (nm, v))
        lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
