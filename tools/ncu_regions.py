"""Per-region breakdown of an ncu source page (SASS view):  python tools/ncu_regions.py gpurun_out/x.ncu-rep [bucket_instrs]
Buckets consecutive SASS instructions and prints, per bucket, executed warp instructions, stall samples and the top stall reasons,
so that the phases of a long kernel (descents / tensor-memory steps / register subtrees / output) can be told apart."""
import csv, io, subprocess, sys
rep = sys.argv[1]
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 128
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
data = rows[2:]
tot_s = sum(int(r[col["# Samples"]] or 0) for r in data)
tot_i = sum(int(r[col["Instructions Executed"]] or 0) for r in data)
print("total samples %d, warp instructions %d, SASS instructions %d" % (tot_s, tot_i, len(data)))
for b0 in range(0, len(data), bucket):
    blk = data[b0:b0 + bucket]
    s = sum(int(r[col["# Samples"]] or 0) for r in blk)
    i = sum(int(r[col["Instructions Executed"]] or 0) for r in blk)
    if s < tot_s * 0.004 and i < tot_i * 0.004:
        continue
    st = sorted(((sum(int(r[col[h]] or 0) for r in blk), h[6:]) for h in stalls), reverse=True)[:4]
    ops = {}
    for r in blk:
        op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
        if op.startswith("@"):
            op = r[col["Source"]].split()[1]
        ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + int(r[col["Instructions Executed"]] or 0)
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:4]
    print("[%5d..%5d] instr %5.1f%%  samples %5.1f%%  samples/instr %.2f | %s | %s" % (
        b0, b0 + len(blk), 100.0 * i / tot_i, 100.0 * s / tot_s, (s / tot_s) / max(i / tot_i, 1e-9),
        " ".join("%s %.0f%%" % (n, 100.0 * v / max(s, 1)) for v, n in st), " ".join("%s" % k for k, v in top)))
