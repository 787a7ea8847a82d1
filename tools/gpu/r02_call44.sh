#!/bin/bash
# r02 call 44: Linear+skip GEMM with 12 epilogue warps; normalise-and-cast with bulk row copies: timing + parity + bench A/B.
mkdir -p gpurun_out
LOG=gpurun_out/r02c44_fc12.log
: > $LOG
for ax in time freq; do timeout 300 python tools/prof_gemm.py --which fc --axis $ax --reps 3 2>&1 | tail -1 >> $LOG; done
cat $LOG
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02c44_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02c44_pytest.log
for pb in 1 0; do
BSRNN_PACK_BULK=$pb timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c44_bench_pb$pb.json 2> gpurun_out/r02c44_bench_pb$pb.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02c44_bench_pb$pb.json').read().strip().splitlines()[-1])
print('bulk=$pb', round(d['ms_per_step'],1), round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], {k:round(v,1) for k,v in d['roofline']['regions_ms_per_step'].items()})
PY
done
