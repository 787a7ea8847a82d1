#!/bin/bash
# r02 call 26: fused layer kernel at the FlowSE width (H = 768): unit parity, FlowSE model tests, config 4 bench A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "fused768 or fused_vs_torch" > gpurun_out/r02c26_pytest_unit.log 2>&1; echo "unit rc=$?"; tail -4 gpurun_out/r02c26_pytest_unit.log
timeout 1200 python -m pytest tests -m gpu -q -x -k "flowse" > gpurun_out/r02c26_pytest_flowse.log 2>&1; echo "flowse rc=$?"; tail -4 gpurun_out/r02c26_pytest_flowse.log
timeout 900 python bench.py --config 4 --steps 2 --warmup 1 --no-cpu-baseline --no-library-baseline > gpurun_out/r02c26_cfg4_fused.json 2> gpurun_out/r02c26_cfg4_fused.err; echo "cfg4 fused rc=$?"
BSRNN_FLOWSE_FUSED=0 timeout 900 python bench.py --config 4 --steps 2 --warmup 1 --no-cpu-baseline --no-library-baseline > gpurun_out/r02c26_cfg4_steps.json 2> gpurun_out/r02c26_cfg4_steps.err; echo "cfg4 steps rc=$?"
python - <<'PY'
import json
for n in ("fused","steps"):
    try:
        d=json.loads(open(f'gpurun_out/r02c26_cfg4_{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['ms_per_step'],1), round(d['value'],2), d['roofline']['frac'] if d.get('roofline') else None, d['clocks']['sm_mhz'], d.get('gpu_launches'))
    except Exception as e:
        print(n,'ERR',e); print(open(f'gpurun_out/r02c26_cfg4_{n}.err').read()[-1500:])
PY
