#!/bin/bash
# GPU call 8: is the publish fence slow because of the SM's outstanding generic stores?  flags 4 = h not stored to
# global, 8 = no fence before the remote arrives (timing only; results are garbage).
mkdir -p gpurun_out
LOG=gpurun_out/call8_lstm_fence.log
: > $LOG
P="timeout 120 python tools/prof_lstm.py --B 64 --T 201 --K 34 --reps 2 --axis time --variant 0 --trace --trace-cid 1"
for f in 0 4 8 12 14; do
  $P --slots 3 --flags $f >> $LOG 2>&1
done
for f in 4 8 12; do
  $P --slots 1 --flags $f >> $LOG 2>&1
done
$P --slots 2 --flags 12 >> $LOG 2>&1
grep -vE "^ +[0-9]+ +0 " $LOG | grep -vE "^step|slot-0 chain" | tail -120
