#!/bin/bash
# GPU call 46: v8 without the per-thread L2 prefetch of the next step's gates_x (register refill runs an item ahead).
mkdir -p gpurun_out
LOG=gpurun_out/call46_lstm.log; : > $LOG
P="timeout 120 python tools/prof_lstm.py"
$P --B 40 --T 60 --K 34 --axis time --slots 3 --maxcl 2 --check --reps 1 >> $LOG 2>&1 || echo "FAILED multi-group time" >> $LOG
$P --B 3 --T 300 --K 34 --axis freq --slots 3 --check --reps 1 >> $LOG 2>&1 || echo "FAILED freq" >> $LOG
for ax in time freq; do
  $P --B 64 --T 1001 --K 34 --axis $ax --slots 3 --reps 4 >> $LOG 2>&1
done
grep -E "CHECK|FAILED|ms,|rror" $LOG
