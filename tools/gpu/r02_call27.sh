#!/bin/bash
# r02 call 27: H = 768 fused kernel alone at config-4 sizes (probe build): slots sweep + per-role cycles; regions of one evaluation.
mkdir -p gpurun_out
LOG=gpurun_out/r02c27_fused768.log
: > $LOG
run() { timeout 300 python tools/prof_lstm768.py "$@" >> $LOG 2>&1 || echo "FAILED rc=$? : $*" >> $LOG; }
run --R 1536 --steps 1251 --slots 1
run --R 1536 --steps 1251 --slots 2 --trace
run --R 1536 --steps 1251 --slots 3
run --R 40032 --steps 48 --slots 3 --trace
run --R 40032 --steps 48 --slots 2
BSRNN_FLOWSE_REGIONS=1 timeout 600 python tools/bench_flowse.py --batch 32 --nfe 2 --graph >> $LOG 2>&1
grep -v Warning $LOG | tail -60
