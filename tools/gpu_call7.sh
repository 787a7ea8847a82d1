#!/bin/bash
# GPU call 7: where does the multi-slot recurrence lose its time?  whole-launch per-role wait/busy cycles, cluster-count
# and gates_x-traffic experiments at a 5x shorter config-2 shape (B=64, T=201, K=34: same 34 units / 2176 sequences).
mkdir -p gpurun_out
LOG=gpurun_out/call7_lstm_probe.log
: > $LOG
P="timeout 120 python tools/prof_lstm.py --B 64 --T 201 --K 34 --reps 2"
$P --axis time --slots 3 --variant 0 --trace --trace-cid 1 >> $LOG 2>&1
$P --axis time --slots 3 --variant 1 --trace --trace-cid 1 >> $LOG 2>&1
$P --axis time --slots 3 --variant 0 --maxcl 1 --trace --trace-cid 0 >> $LOG 2>&1
$P --axis time --slots 3 --variant 0 --maxcl 4 --trace --trace-cid 1 >> $LOG 2>&1
$P --axis time --slots 3 --variant 0 --maxcl 8 >> $LOG 2>&1
$P --axis time --slots 3 --variant 0 --flags 1 >> $LOG 2>&1
$P --axis time --slots 3 --variant 0 --flags 2 >> $LOG 2>&1
$P --axis time --slots 3 --variant 0 --flags 3 --trace --trace-cid 1 >> $LOG 2>&1
$P --axis time --slots 2 --variant 0 --trace --trace-cid 1 >> $LOG 2>&1
$P --axis time --slots 1 --variant 0 --trace --trace-cid 1 >> $LOG 2>&1
$P --axis freq --slots 1 --variant 0 --trace --trace-cid 1 >> $LOG 2>&1
$P --axis freq --slots 2 --variant 0 --trace --trace-cid 1 >> $LOG 2>&1
$P --axis freq --slots 3 --variant 0 --trace --trace-cid 1 >> $LOG 2>&1
$P --axis freq --slots 3 --variant 0 --flags 3 >> $LOG 2>&1
$P --axis freq --slots 3 --variant 0 --maxcl 4 >> $LOG 2>&1
$P --axis freq --slots 4 --variant 1 --trace --trace-cid 1 >> $LOG 2>&1
grep -vE "^ +[0-9]+ +0 " $LOG | grep -vE "^step|slot-0 chain" | tail -120
