#!/bin/bash
# r02 call 36 (1 GPU): the default bench line (every leg), the reference arm, configs 3 / 4 / 5 with the fused schedule.
mkdir -p gpurun_out
S=$(date +%s)
timeout 1200 python bench.py > gpurun_out/r02c36_bench_cfg2.json 2> gpurun_out/r02c36_bench_cfg2.err; echo "cfg2 rc=$? wall=$(( $(date +%s) - S )) s"; tail -2 gpurun_out/r02c36_bench_cfg2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c36_bench_cfg2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','scaling','dtype')}, d['e2e'], d['roofline']['frac'], d['roofline']['per_axis'], d['roofline']['whole_step'])
print('cpu', d.get('cpu_baseline')); print('lib', d.get('library_baseline')); print('fp32', d.get('fp32_mode')); print(d['clocks'])
print({k:(round(v['achieved']),round(v['frac'],2)) for k,v in d['roofline']['other_kernels'].items()})
print({k:round(v,2) for k,v in d['roofline']['regions_ms_per_step'].items()})
PY
S=$(date +%s)
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | tail -1 | cut -c 1-600; echo "ref wall=$(( $(date +%s) - S )) s"
for c in 3 4 5; do
  S=$(date +%s)
  timeout 900 python bench.py --config $c > gpurun_out/r02c36_bench_cfg$c.json 2> gpurun_out/r02c36_bench_cfg$c.err; echo "cfg$c rc=$? wall=$(( $(date +%s) - S )) s"; tail -2 gpurun_out/r02c36_bench_cfg$c.err; cat gpurun_out/r02c36_bench_cfg$c.json | cut -c 1-1200
done
