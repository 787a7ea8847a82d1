#!/bin/bash
# r02 call 34: full gpu suite + config 2 bench with the automatic fused geometry (7 pairs x 56 units where it splits better).
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02c34_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02c34_pytest.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c34_bench.json 2> gpurun_out/r02c34_bench.err; echo "bench rc=$?"
BSRNN_LSTM_FUSED_GEO=8 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c34_bench_geo8.json 2> gpurun_out/r02c34_bench_geo8.err; echo "bench rc=$?"
python - <<'PY'
import json
for n in ("bench","bench_geo8"):
    d=json.loads(open(f'gpurun_out/r02c34_{n}.json').read().strip().splitlines()[-1])
    print(n, round(d['ms_per_step'],1), round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['gpu_launches']/d['steps'], {k:round(v,1) for k,v in d['roofline']['regions_ms_per_step'].items()})
PY
