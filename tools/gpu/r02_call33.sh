#!/bin/bash
# r02 call 33: 7-pair geometry (56 units per pair, 10 groups) of the fused H = 392 layer kernel: parity + timing vs 8 pairs.
mkdir -p gpurun_out
LOG=gpurun_out/r02c33_fused7.log
: > $LOG
timeout 900 python -m pytest tests -m gpu -q -x -k "fused_vs_torch" > gpurun_out/r02c33_pytest_unit.log 2>&1; echo "unit rc=$?"; tail -3 gpurun_out/r02c33_pytest_unit.log
run() { timeout 300 python tools/prof_lstm.py "$@" >> $LOG 2>&1 || echo "FAILED rc=$? : $*" >> $LOG; }
run --fused --geo 7 --B 64 --T 1001 --K 34 --axis time
run --fused --geo 8 --B 64 --T 1001 --K 34 --axis time
run --fused --geo 7 --B 64 --T 1001 --K 34 --axis freq
run --fused --geo 8 --B 64 --T 1001 --K 34 --axis freq
run --fused --geo 7 --B 64 --T 1001 --K 34 --axis time --slots 3
grep -v Warning $LOG | tail -40
