#!/bin/bash
# GPU call 31 (fine-grained 4 KB h ring) (1 GPU): recurrence v8 (v5 + register-refilled gates, FHADD, pipelined TMEM loads) vs v5 (now with FHADD):
# parity at small shapes, timing + role trace at BASELINE config 2.
mkdir -p gpurun_out
LOG=gpurun_out/call31_lstm.log; : > $LOG
for v in 8 5; do
  P="timeout 120 python tools/prof_lstm.py --ver $v"
  $P --B 12 --T 40 --K 34 --axis time --slots 1 --check --reps 1 >> $LOG 2>&1 || echo "FAILED v$v time slots=1" >> $LOG
  $P --B 10 --T 40 --K 34 --axis time --slots 2 --check --reps 1 >> $LOG 2>&1 || echo "FAILED v$v time slots=2" >> $LOG
  $P --B 3 --T 300 --K 34 --axis freq --slots 3 --check --reps 1 >> $LOG 2>&1 || echo "FAILED v$v freq slots=3" >> $LOG
  $P --B 40 --T 60 --K 34 --axis time --slots 3 --maxcl 2 --check --reps 1 >> $LOG 2>&1 || echo "FAILED v$v multi-group time" >> $LOG
  $P --B 40 --T 60 --K 34 --axis freq --slots 3 --maxcl 2 --check --reps 1 >> $LOG 2>&1 || echo "FAILED v$v multi-group freq" >> $LOG
  for ax in time freq; do
    $P --B 64 --T 1001 --K 34 --axis $ax --slots 3 --reps 3 --trace >> $LOG 2>&1
  done
done
grep -E "CHECK|FAILED|ms,|producer|mma  |epilogue|rror" $LOG
