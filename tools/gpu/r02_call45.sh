#!/bin/bash
# r02 call 45: TMA residual epilogue of the Linear+skip GEMM: equality with the register-staged epilogue, timing, suites, bench.
mkdir -p gpurun_out
LOG=gpurun_out/r02c45_fc_tma.log
: > $LOG
timeout 300 python tools/check_fc_tma.py >> $LOG 2>&1
for ax in time freq; do
  timeout 300 python tools/prof_gemm.py --which fc --axis $ax --reps 3 --tma 2>&1 | tail -1 >> $LOG
  timeout 300 python tools/prof_gemm.py --which fc --axis $ax --reps 3 2>&1 | tail -1 >> $LOG
done
cat $LOG
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02c45_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02c45_pytest.log
for e in tma ldst; do
BSRNN_FC_EPI=$e timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c45_bench_$e.json 2> gpurun_out/r02c45_bench_$e.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02c45_bench_$e.json').read().strip().splitlines()[-1])
print('$e', round(d['ms_per_step'],1), round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], {k:round(v,1) for k,v in d['roofline']['regions_ms_per_step'].items()})
PY
done
