#!/bin/bash
# r02 call 29: x-tile L2 prefetch one step ahead in the fused kernels (both widths): timing alone + parity units.
mkdir -p gpurun_out
LOG=gpurun_out/r02c29_prefetch.log
: > $LOG
timeout 900 python -m pytest tests -m gpu -q -x -k "fused768 or fused_vs_torch" > gpurun_out/r02c29_pytest_unit.log 2>&1; echo "unit rc=$?"; tail -3 gpurun_out/r02c29_pytest_unit.log
run() { timeout 300 python tools/prof_lstm768.py "$@" >> $LOG 2>&1 || echo "FAILED rc=$? : $*" >> $LOG; }
run --R 1536 --steps 1251 --slots 2
run --R 40032 --steps 48 --slots 3
run2() { timeout 300 python tools/prof_lstm.py "$@" >> $LOG 2>&1 || echo "FAILED rc=$? : $*" >> $LOG; }
run2 --fused --B 64 --T 1001 --K 34 --axis time
run2 --fused --B 64 --T 1001 --K 34 --axis freq
grep -v Warning $LOG | tail -40
