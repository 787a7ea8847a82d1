"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per (kernel, grid) count and total time.
usage: python tools/summarize_launches.py launches.csv [skip_first_n] [title]"""
import csv, sys, collections, io
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = list(csv.reader(io.StringIO("".join(rows))))
hdr, r = r[0], r[1:]
ik, ig, iv = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
r = r[skip:]
agg = collections.OrderedDict()
for x in r:
    k = (x[ik].split("(")[0][:60], x[ig])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += float(x[iv].replace(",", "")) / 1e6
tot = sum(a[1] for a in agg.values())
if len(sys.argv) > 3: print(sys.argv[3])
for (k, g), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} grid={g:>16s} n={n:3d} total={t:8.3f} ms {100*t/tot:6.1f}%")
print(f"total {tot:.3f} ms over {len(r)} launches")
