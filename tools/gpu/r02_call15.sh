#!/bin/bash
# r02 call 15: STFT / iSTFT with two frames per complex FFT, STFT-fused band statistics, dedicated band-split kernel:
# full gpu suite, then bench with per-kernel timings of the small kernels (launch list).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02c15_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02c15_pytest.log
KREG='regex:^(lstm_|gemm_|norm_cast|istft|stft|band_|gn_)'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 207 -c 210 --csv --log-file gpurun_out/r02c15_ncu_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c15_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import csv,collections,re
rows=list(csv.reader(open('gpurun_out/r02c15_ncu_launches_bench.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; data=rows[hdr+1:]
ki=H.index('Kernel Name'); vi=H.index('Metric Value'); gi=H.index('Grid Size')
agg=collections.OrderedDict()
for r in data:
    if len(r)<=vi: continue
    k=re.sub(r'\(.*','',r[ki]); a=agg.setdefault((k,r[gi]),[0,0.0]); a[0]+=1; a[1]+=float(r[vi].replace(',',''))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    if 'lstm' in k[0] or 'gemm_tc_kernel<4' in k[0] or 'gemm_tc_kernel<1' in k[0]: continue
    print(f"{k[0][:50]:50s} grid={k[1]:>16s} n={v[0]:3d} total={v[1]/1e6:7.3f} ms")
PY
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c15_bench.json 2> gpurun_out/r02c15_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02c15_bench.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],1), round(d['value']), round(d['e2e']['value']), d['roofline']['frac'], d['clocks']['sm_mhz'], d['gpu_launches']/d['steps'])"
