#!/bin/bash
# r02 call 28: H = 768 fused kernel with 4-CTA clusters (x / h stages multicast to the two CTAs of a parity): parity + timing.
mkdir -p gpurun_out
LOG=gpurun_out/r02c28_fused768_mc.log
: > $LOG
timeout 900 python -m pytest tests -m gpu -q -x -k "fused768 or fused_vs_torch" > gpurun_out/r02c28_pytest_unit.log 2>&1; echo "unit rc=$?"; tail -3 gpurun_out/r02c28_pytest_unit.log
run() { timeout 300 python tools/prof_lstm768.py "$@" >> $LOG 2>&1 || echo "FAILED rc=$? : $*" >> $LOG; }
run --R 1536 --steps 1251 --slots 2 --trace
run --R 1536 --steps 1251 --slots 3
run --R 40032 --steps 48 --slots 3 --trace
grep -v Warning $LOG | tail -40
