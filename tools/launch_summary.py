"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum[,dram__bytes_*] --csv) by kernel.
   python tools/launch_summary.py gpurun_out/launches.csv"""
import csv
import collections
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.defaultdict(dict)
    names = {}
    for r in rows[1:]:
        per[r[iid]][r[im]] = float(r[iv].replace(",", ""))
        names[r[iid]] = r[ik]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, []])
    for i, m in per.items():
        a = agg[names[i].split("(")[0]]
        a[0] += 1
        t = m.get("gpu__time_duration.sum", 0.0)
        a[1] += t
        a[2] += m.get("dram__bytes_read.sum", 0.0)
        a[3] += m.get("dram__bytes_write.sum", 0.0)
        a[4].append((t, m.get("dram__bytes_read.sum", 0.0), m.get("dram__bytes_write.sum", 0.0)))
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total device time (ns->ms) | share | longest launch (ms) | DRAM read / write of the longest (GB) |")
    print("|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        big = max(a[4])
        print("| `%s` | %d | %.2f | %.1f %% | %.3f | %.2f / %.3f |" % (k, a[0], a[1] / 1e6, 100 * a[1] / tot, big[0] / 1e6, big[1] / 1e9, big[2] / 1e9))


if __name__ == "__main__":
    main()
