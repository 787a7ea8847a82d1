#!/bin/bash
# GPU call 20: v5 recurrence epilogue (all 12 warps per accumulator item): parity, timing vs v4, probe, ncu; then tests + bench.
mkdir -p gpurun_out
LOG=gpurun_out/call20_lstm_v5.log
: > $LOG
P="timeout 120 python tools/prof_lstm.py"
for sl in 1 2 3; do
  $P --B 12 --T 40 --K 34 --axis time --slots $sl --check --reps 1 >> $LOG 2>&1 || echo "FAILED time slots=$sl rc=$?" >> $LOG
  $P --B 3 --T 300 --K 34 --axis freq --slots $sl --check --reps 1 >> $LOG 2>&1 || echo "FAILED freq slots=$sl rc=$?" >> $LOG
done
$P --B 40 --T 60 --K 34 --axis time --slots 3 --maxcl 3 --check --reps 1 >> $LOG 2>&1 || echo "FAILED multi-group time" >> $LOG
$P --B 40 --T 60 --K 34 --axis freq --slots 3 --maxcl 2 --check --reps 1 >> $LOG 2>&1 || echo "FAILED multi-group freq" >> $LOG
for ax in time freq; do
  BSRNN_LSTM_V4=1 $P --B 64 --T 1001 --K 34 --axis $ax --slots 3 --reps 3 >> $LOG 2>&1
  $P --B 64 --T 1001 --K 34 --axis $ax --slots 3 --reps 3 --trace >> $LOG 2>&1
  $P --B 64 --T 1001 --K 34 --axis $ax --slots 2 --reps 2 >> $LOG 2>&1
done
grep -E "CHECK|FAILED|ms,|cycles|producer|mma  |epilogue|rror" $LOG | tail -50
if grep -q "FAIL" $LOG; then export BSRNN_LSTM_V4=1; echo "v5 FAILED -> evidence on v4"; fi
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/call20_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/call20_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/call20_bench.json 2> gpurun_out/call20_bench.err; echo "bench rc=$?"; cat gpurun_out/call20_bench.json; tail -3 gpurun_out/call20_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tc -s 1 -c 1 -o gpurun_out/call20_lstm_v5_full python tools/prof_lstm.py --B 64 --T 1001 --K 34 --axis time --slots 3 --reps 2 > gpurun_out/call20_ncu_lstm.log 2>&1; echo "ncu lstm rc=$?"
