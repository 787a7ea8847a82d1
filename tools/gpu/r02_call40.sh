#!/bin/bash
# r02 call 40: 14-pair small-batch geometry: parity, timing at small batches, bench at B = 8 / 16 / 32, config 3.
mkdir -p gpurun_out
LOG=gpurun_out/r02c40_fused14.log
: > $LOG
timeout 900 python -m pytest tests -m gpu -q -x -k "fused_vs_torch or schedules_agree" > gpurun_out/r02c40_pytest_unit.log 2>&1; echo "unit rc=$?"; tail -3 gpurun_out/r02c40_pytest_unit.log
run() { timeout 300 python tools/prof_lstm.py "$@" >> $LOG 2>&1 || echo "FAILED rc=$? : $*" >> $LOG; }
for b in 8 16 32; do
  run --fused --geo 14 --B $b --T 1001 --K 34 --axis time --reps 2
  run --fused --geo 7 --B $b --T 1001 --K 34 --axis time --reps 2
done
run --fused --geo 14 --B 8 --T 1001 --K 34 --axis freq --reps 2
run --fused --geo 7 --B 8 --T 1001 --K 34 --axis freq --reps 2
grep -v Warning $LOG | grep -v "co-resident" | tail -40
for b in 8 16 32; do
  timeout 600 python bench.py --batch $b --steps 4 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c40_bench_b$b.json 2> gpurun_out/r02c40_bench_b$b.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02c40_bench_b$b.json').read().strip().splitlines()[-1])
print('B=$b', round(d['ms_per_step'],1), round(d['value']), round(d['e2e']['value']), d['clocks']['sm_mhz'], {k:round(v,1) for k,v in d['roofline']['regions_ms_per_step'].items()})
PY
done
timeout 600 python bench.py --config 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3', round(d['value']), round(d['e2e']['value']), {k:round(v['audio_s_per_s']) for k,v in d['per_rate'].items()})"
