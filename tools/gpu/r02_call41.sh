#!/bin/bash
# r02 call 41: streaming GEMMs with 8 stages of 4 k-cores instead of 4 stages of 8 (Linear 4N->N): parity + A/B.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02c41_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02c41_pytest.log
for ks in 4 8; do
BSRNN_GEMM_KS=$ks timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c41_bench_ks$ks.json 2> gpurun_out/r02c41_bench_ks$ks.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02c41_bench_ks$ks.json').read().strip().splitlines()[-1])
print('ks=$ks', round(d['ms_per_step'],1), round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], {k:round(v,1) for k,v in d['roofline']['regions_ms_per_step'].items()})
PY
done
for ks in 4 8; do
BSRNN_GEMM_KS=$ks BSRNN_FLOWSE_REGIONS=1 timeout 600 python tools/bench_flowse.py --batch 32 --nfe 2 --graph 2>&1 | grep -v Warn | tail -2
done
