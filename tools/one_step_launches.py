"""Per-kernel totals of the LAST forward in an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py.
usage: python tools/one_step_launches.py launches.csv [...]"""
import csv, io, sys, collections


def one_forward(path):
    rows = [l for l in open(path) if l.startswith('"')]
    r = list(csv.reader(io.StringIO("".join(rows)))); hdr, r = r[0], r[1:]
    ik, iv, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    starts = [i for i, x in enumerate(r) if 'stft_kernel' in x[ik] and 'istft' not in x[ik]]
    return [(x[ik].split("(")[0].replace("void ", "")[:40], x[ig], float(x[iv].replace(",", "")) / 1e3) for x in r[starts[-1]:]]


for f in sys.argv[1:]:
    o = one_forward(f)
    agg = collections.OrderedDict()
    for k, g, t in o:
        if k.startswith('gemm_tc_kernel<8'):
            k += ' (band split)' if t < 600 else ' (Linear+skip)'
        a = agg.setdefault((k, g), [0, 0.0]); a[0] += 1; a[1] += t
    tot = sum(t for _, _, t in o)
    print(f"{f}: one step = {tot / 1e3:.3f} ms of kernel time over {len(o)} launches (kernels alone, un-capped clocks)")
    for (k, g), (n, t) in agg.items():
        print(f"   {k:46s} {g:16s} n={n:3d} {t / 1e3:8.3f} ms {100 * t / tot:5.1f}%  avg {t / n:8.1f} us")
