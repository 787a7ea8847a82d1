#!/bin/bash
# r02 call 25: mask decoder on two graph branches: parity tests + bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "parity or reference or fullsize or inference" > gpurun_out/r02c25_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02c25_pytest.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c25_bench.json 2> gpurun_out/r02c25_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c25_bench.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],1), round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['gpu_launches']/d['steps'], {k:round(v,1) for k,v in d['roofline']['regions_ms_per_step'].items()})
PY
