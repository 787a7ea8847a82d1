#!/bin/bash
# r02 call 2 (1 GPU): flag-group recurrence: numerics vs torch, timing against the cluster schedule at config-2 size,
# strong-scaling sizes; then the remaining reference tests and the bench in both schedules.
mkdir -p gpurun_out
for ax in time freq; do
  timeout 120 python tools/prof_lstm.py --B 3 --T 61 --K 34 --axis $ax --flag --check --reps 1 2>&1 | grep -E "CHECK|Error|error" | tail -2
done
timeout 120 python tools/prof_lstm.py --B 20 --T 301 --K 34 --axis time --flag --check --reps 1 2>&1 | grep -E "CHECK|Error|error" | tail -2
for args in "--B 64 --axis time" "--B 64 --axis time --slots 3" "--B 64 --axis freq" "--B 64 --axis freq --slots 2" "--B 8 --axis time" "--B 8 --axis freq" "--B 16 --axis time" "--B 32 --axis time"; do
  timeout 120 python tools/prof_lstm.py --T 1001 --K 34 $args --flag --reps 3 2>&1 | tail -1
done
timeout 600 python -m pytest tests/test_gpu_reference.py tests/test_gpu_training.py -m gpu -q -s > gpurun_out/r02c02_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "rel_l2|passed|failed|Error|error|band-limited" gpurun_out/r02c02_pytest.log | tail -30
for sched in cluster flag; do
  BSRNN_LSTM_SCHED=$sched timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02c02_bench_$sched.json 2> gpurun_out/r02c02_bench_$sched.err; echo "bench $sched rc=$?"
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r02c02_bench_$sched.json').read().strip().splitlines()[-1])
print('$sched', d['value'], d['ms_per_step'], d.get('regions_ms_per_step'), d['roofline']['frac'], d['e2e']['value'], d.get('clocks'))"
done
