#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r02c21_fused_probe.log
: > $LOG
run() { timeout 300 python tools/prof_lstm.py "$@" >> $LOG 2>&1 || echo "FAILED rc=$? : $*" >> $LOG; }
run --fused --check --reps 1 --B 10 --T 40 --K 34 --axis time --slots 2
run --fused --check --reps 1 --B 3 --T 300 --K 34 --axis freq --slots 3
run --fused --trace --reps 2 --B 12 --T 40 --K 34 --axis time --slots 1
run --fused --trace --reps 2 --B 64 --T 1001 --K 34 --axis freq --slots 3
run --fused --trace --reps 2 --B 64 --T 1001 --K 34 --axis time --slots 3
run --fused --reps 2 --B 64 --T 1001 --K 34 --axis time --slots 2
grep -v Warning $LOG | tail -50
