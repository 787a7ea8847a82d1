#!/bin/bash
# GPU call 45 (1 GPU): recurrence with the debug probes compiled out: parity at small shapes, timing, pytest, bench.
mkdir -p gpurun_out
LOG=gpurun_out/call45_lstm.log; : > $LOG
P="timeout 120 python tools/prof_lstm.py"
$P --B 12 --T 40 --K 34 --axis time --slots 1 --check --reps 1 >> $LOG 2>&1 || echo "FAILED time slots=1" >> $LOG
$P --B 10 --T 40 --K 34 --axis time --slots 2 --check --reps 1 >> $LOG 2>&1 || echo "FAILED time slots=2" >> $LOG
$P --B 3 --T 300 --K 34 --axis freq --slots 3 --check --reps 1 >> $LOG 2>&1 || echo "FAILED freq slots=3" >> $LOG
$P --B 40 --T 60 --K 34 --axis time --slots 3 --maxcl 2 --check --reps 1 >> $LOG 2>&1 || echo "FAILED multi-group time" >> $LOG
$P --B 40 --T 60 --K 34 --axis freq --slots 3 --maxcl 2 --check --reps 1 >> $LOG 2>&1 || echo "FAILED multi-group freq" >> $LOG
for ax in time freq; do
  $P --B 64 --T 1001 --K 34 --axis $ax --slots 3 --reps 4 >> $LOG 2>&1
  $P --ver 5 --B 64 --T 1001 --K 34 --axis $ax --slots 3 --reps 3 >> $LOG 2>&1
done
grep -E "CHECK|FAILED|ms,|rror" $LOG
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/call45_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/call45_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/call45_bench.json 2> gpurun_out/call45_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/call45_bench.json"))
print(round(d["ms_per_step"],1), "e2e", round(d["per_rank"]["e2e_ms_per_step"][0],1), d["clocks"], {k:round(x,1) for k,x in d["roofline"]["regions_ms_per_step"].items()}, "frac", round(d["roofline"]["frac"],3))
PY
