#!/bin/bash
# r02 call 46: TMA residual epilogue everywhere (FlowSE condition_fc / Linear blocks, training forward): suites + configs 4, 5, 3.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02c46_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02c46_pytest.log
for c in 4 5 3; do
  timeout 900 python bench.py --config $c --no-cpu-baseline --no-library-baseline > gpurun_out/r02c46_bench_cfg$c.json 2> gpurun_out/r02c46_bench_cfg$c.err; echo "cfg$c rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02c46_bench_cfg$c.json').read().strip().splitlines()[-1])
print('cfg$c', round(d['ms_per_step'],1), round(d['value'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'], (d.get('roofline') or {}).get('frac'))
PY
done
BSRNN_FLOWSE_REGIONS=1 timeout 600 python tools/bench_flowse.py --batch 32 --nfe 2 --graph 2>&1 | grep -v Warn | tail -2
