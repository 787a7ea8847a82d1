#!/bin/bash
# r02 call 17: fused BLSTM layer (CTA pairs + flag groups + x*W_ih inside the recurrence): parity on small shapes, then
# timing at BASELINE config 2 sizes against the unfused flag schedule.
mkdir -p gpurun_out
LOG=gpurun_out/r02c17_fused.log
: > $LOG
run() { timeout 300 python tools/prof_lstm.py "$@" >> $LOG 2>&1 || echo "FAILED rc=$? : $*" >> $LOG; }
run --fused --check --B 12 --T 40 --K 34 --axis time --slots 1
run --fused --check --B 10 --T 40 --K 34 --axis time --slots 2
run --fused --check --B 3 --T 300 --K 34 --axis freq --slots 3
run --fused --check --B 40 --T 60 --K 34 --axis freq --slots 0
run --fused --check --B 40 --T 60 --K 34 --axis time --slots 0 --maxcl 2
run --flag --B 64 --T 1001 --K 34 --axis time
run --fused --B 64 --T 1001 --K 34 --axis time --slots 3
run --fused --B 64 --T 1001 --K 34 --axis time --slots 2
run --flag --B 64 --T 1001 --K 34 --axis freq
run --fused --B 64 --T 1001 --K 34 --axis freq --slots 3
run --fused --B 64 --T 1001 --K 34 --axis freq --slots 2
cat $LOG | grep -v Warning | tail -60
