#!/bin/bash
# GPU call 6: v3 multi-slot recurrence: parity vs torch per (slots, variant), timing at config-2 shapes, then the
# evidence set on the current default path (pytest -m gpu, bench, ncu launch list).
mkdir -p gpurun_out
LOG=gpurun_out/call6_lstm_v3.log
: > $LOG
for cfg in "1 0" "2 0" "3 0" "3 1" "4 1"; do
  set -- $cfg
  timeout 120 python tools/prof_lstm.py --B 12 --T 40 --K 34 --axis time --slots $1 --variant $2 --check --reps 1 >> $LOG 2>&1 || echo "FAILED time slots=$1 variant=$2 rc=$?" >> $LOG
  timeout 120 python tools/prof_lstm.py --B 3 --T 300 --K 34 --axis freq --slots $1 --variant $2 --check --reps 1 >> $LOG 2>&1 || echo "FAILED freq slots=$1 variant=$2 rc=$?" >> $LOG
done
timeout 120 python tools/prof_lstm.py --B 64 --T 1001 --K 34 --axis time --v2 --reps 2 >> $LOG 2>&1
timeout 120 python tools/prof_lstm.py --B 64 --T 1001 --K 34 --axis freq --v2 --reps 2 >> $LOG 2>&1
for cfg in "1 0" "2 0" "3 0" "3 1" "4 1"; do
  set -- $cfg
  timeout 120 python tools/prof_lstm.py --B 64 --T 1001 --K 34 --axis time --slots $1 --variant $2 --reps 2 >> $LOG 2>&1 || echo "FAILED big time $cfg" >> $LOG
  timeout 120 python tools/prof_lstm.py --B 64 --T 1001 --K 34 --axis freq --slots $1 --variant $2 --reps 2 >> $LOG 2>&1 || echo "FAILED big freq $cfg" >> $LOG
done
timeout 120 python tools/prof_lstm.py --B 64 --T 1001 --K 34 --axis time --slots 3 --variant 0 --reps 1 --trace >> $LOG 2>&1
grep -v "^$" $LOG | grep -E "CHECK|FAILED|ms,|round period|Error|error" | tail -60
# ---- evidence set on the best working schedule
if grep -q "\[v3 slots=3 variant=0\] CHECK time.*OK" $LOG && grep -q "\[v3 slots=3 variant=0\] CHECK freq.*OK" $LOG; then
  echo "v3 slots=3 ok: default path" | tee -a $LOG
else
  export BSRNN_LSTM_V2=1; echo "v3 FAILED: falling back to v2 for the evidence set" | tee -a $LOG
fi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/call6_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/call6_pytest_gpu.log; tail -3 gpurun_out/call6_pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/call6_bench.json 2> gpurun_out/call6_bench.err; echo "bench rc=$?"; cat gpurun_out/call6_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/call6_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/call6_ncu_bench.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/call6_launches.csv
