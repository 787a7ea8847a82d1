#!/bin/bash
# r02 call 3 (1 GPU): remaining reference / training / metric tests, then the new bench.py in every config.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reference.py tests/test_gpu_training.py -m gpu -q -s > gpurun_out/r02c03_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "FlowSE N=384|ESTOI|passed|failed|Error|error" gpurun_out/r02c03_pytest.log | tail -30
timeout 900 python bench.py > gpurun_out/r02c03_bench_cfg2.json 2> gpurun_out/r02c03_bench_cfg2.err; echo "cfg2 rc=$?"; tail -3 gpurun_out/r02c03_bench_cfg2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c03_bench_cfg2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','scaling','dtype')}, d['e2e'], d['roofline']['frac'], d['roofline']['per_axis'], d['roofline']['whole_step'])
print('cpu', d.get('cpu_baseline')); print('lib', d.get('library_baseline')); print('fp32', d.get('fp32_mode')); print(d['clocks'])
print({k:(round(v['achieved']),round(v['frac'],2)) for k,v in d['roofline']['other_kernels'].items()})
print({k:round(v,2) for k,v in d['roofline']['regions_ms_per_step'].items()})
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | tail -1
for c in 3 4 5; do
  timeout 900 python bench.py --config $c > gpurun_out/r02c03_bench_cfg$c.json 2> gpurun_out/r02c03_bench_cfg$c.err; echo "cfg$c rc=$?"; tail -2 gpurun_out/r02c03_bench_cfg$c.err; cat gpurun_out/r02c03_bench_cfg$c.json | cut -c 1-1500
done
