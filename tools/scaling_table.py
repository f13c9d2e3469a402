"""Scaling table from the bench lines of N = 1, 2, 4, 8 GPUs:  python tools/scaling_table.py gpurun_out/r4f_bench{N}.json > profiles/r02_scaling.md"""
import json
import sys

pat = sys.argv[1]
rows = {}
for N in (1, 2, 4, 8):
    try:
        rows[N] = json.loads(open(pat.replace("{N}", str(N))).read().strip().splitlines()[-1])
    except OSError:
        pass
b = rows[1]
print("# Scaling 1 -> 8 B200 (one box), `bench.py --gpus N --steps 20 --warmup 5` under the driver's torchrun command\n")
print("Every column is the whole job's rate (max over ranks of the device / wall time).  `x` = ratio to the 1-GPU line.\n")
print("| N | SC n=1024 decode, info Gbit/s (`value`, weak) | x | BLER sweep SCL-32 n=2048 (configs[3]), wall ms | x | decoder launches / items queued / counted | BLER sweep SC n=1024, wall ms | x | SCL-8+CRC11 n=1024 Gbit/s | x | e2e module Gbit/s | e2e packed C ABI Gbit/s |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for N, d in rows.items():
    s, s1 = d["sweep"], b["sweep"]
    a, c = s["scl32_n2048"], s["sc_n1024"]
    print("| %d | %.1f | %.2f | %.1f | **%.2f** | %s / %s / %s | %.1f | **%.2f** | %.2f | %.2f | %.2f | %.2f |" % (
        N, d["value"], d["value"] / b["value"], a["wall_ms"], s1["scl32_n2048"]["wall_ms"] / a["wall_ms"],
        a.get("decoder_launches"), a["iterations_queued"], a["iterations_counted"], c["wall_ms"], s1["sc_n1024"]["wall_ms"] / c["wall_ms"],
        d["scl8"]["value"], d["scl8"]["value"] / b["scl8"]["value"], d["e2e"]["value"], d["e2e"]["packed_c_abi"]["value"]))
print("\nBLER sweeps are STRONG scaling: the per-iteration batch (2^16 resp. 2^20 codewords) is split over the ranks, every decoder launch is "
      "followed by one NCCL all-reduce of its items' 4 counters and the on-device stop rules (`my_sn/sim.py::sim_ber_device`).  Per-item split (us, rank 0):\n")
print("| N | sweep | front end | decode | count | all-reduce | control | device total | host wall |")
print("|---|---|---|---|---|---|---|---|---|")
for N, d in rows.items():
    for key in ("scl32_n2048", "sc_n1024"):
        sp = d["sweep"][key]["split_us_per_iteration"]
        print("| %d | %s | %.0f | %.0f | %.0f | %.1f | %.1f | %.0f | %.0f |" % (N, key, sp["front_end"], sp["decode"], sp["count"], sp["all_reduce"],
                                                                            sp["control"], sp["device_total"], sp["host_wall"]))
print("\nThe end-to-end legs (host buffers, H2D + D2H inside the timed region) stop scaling at the box's host-to-device path: "
      "8 ranks x 4 GiB of fp32 logits per step.  `clocks` of every line: %s" % ", ".join(
          "N=%d %s MHz %s" % (N, d["clocks"]["sm_mhz"], d["clocks"]["reasons"]) for N, d in rows.items()))
