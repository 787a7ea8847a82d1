#!/bin/bash
# r02 call 22: fused BLSTM layer kernel as the default schedule: gpu suite, then bench config 2 (fused) and A/B (unfused).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02c22_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02c22_pytest.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c22_bench_fused.json 2> gpurun_out/r02c22_bench_fused.err; echo "bench rc=$?"
BSRNN_LSTM_FUSED=none timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c22_bench_unfused.json 2> gpurun_out/r02c22_bench_unfused.err; echo "bench rc=$?"
BSRNN_LSTM_FUSED=freq timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c22_bench_fusedfreq.json 2> gpurun_out/r02c22_bench_fusedfreq.err; echo "bench rc=$?"
python - <<'PY'
import json
for n in ("fused","unfused","fusedfreq"):
    try:
        d=json.loads(open(f'gpurun_out/r02c22_bench_{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['ms_per_step'],1), round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['gpu_launches']/d['steps'], {k:round(v,1) for k,v in d['roofline']['regions_ms_per_step'].items()})
    except Exception as e:
        print(n, 'ERR', e)
PY
