#!/bin/bash
# r02 call 32: GradDecoder GEMMs on tcgen05 (TANH_F32 epilogue), async normalise-and-cast at N = 384 / 768: tests + config 4.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "flowse or fused768 or reference or inference" > gpurun_out/r02c32_pytest_flowse.log 2>&1; echo "flowse rc=$?"; tail -4 gpurun_out/r02c32_pytest_flowse.log
timeout 900 python bench.py --config 4 --steps 2 --warmup 1 --no-cpu-baseline --no-library-baseline > gpurun_out/r02c32_cfg4.json 2> gpurun_out/r02c32_cfg4.err; echo "cfg4 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02c32_cfg4.json').read().strip().splitlines()[-1])
    print(round(d['ms_per_step'],1), round(d['value'],2), d['roofline']['frac'], d['clocks']['sm_mhz'], d.get('gpu_launches'))
except Exception as e:
    print('ERR',e); print(open('gpurun_out/r02c32_cfg4.err').read()[-1500:])
PY
BSRNN_FLOWSE_REGIONS=1 timeout 600 python tools/bench_flowse.py --batch 32 --nfe 2 --graph 2>&1 | grep -v Warn | tail -3
