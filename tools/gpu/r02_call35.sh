#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r02c35_fused7_probe.log
: > $LOG
run() { timeout 300 python tools/prof_lstm.py "$@" >> $LOG 2>&1 || echo "FAILED rc=$? : $*" >> $LOG; }
run --fused --geo 7 --trace --reps 2 --B 64 --T 1001 --K 34 --axis time
run --fused --geo 7 --trace --reps 2 --B 64 --T 1001 --K 34 --axis freq
run --fused --geo 7 --trace --reps 2 --B 8 --T 1001 --K 34 --axis time
grep -v Warning $LOG | tail -40
