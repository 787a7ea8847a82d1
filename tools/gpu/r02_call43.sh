#!/bin/bash
# r02 call 43: what paces the Linear+skip GEMM: A/B with parts of its residual epilogue switched off.
mkdir -p gpurun_out
LOG=gpurun_out/r02c43_fc_ab.log
: > $LOG
for dbg in 0 4 8 16 24; do
  echo "=== BSRNN_GEMM_DEBUG=$dbg" >> $LOG
  for ax in time freq; do
    BSRNN_GEMM_DEBUG=$dbg timeout 300 python tools/prof_gemm.py --which fc --axis $ax --reps 3 2>&1 | tail -1 >> $LOG
  done
done
echo "=== BSRNN_GEMM_PFDIST=1" >> $LOG
BSRNN_GEMM_PFDIST=1 timeout 300 python tools/prof_gemm.py --which fc --axis time --reps 3 2>&1 | tail -1 >> $LOG
echo "=== BSRNN_GEMM_STAGES=4 (default is 4 anyway) / KS=4" >> $LOG
BSRNN_GEMM_KS=4 timeout 300 python tools/prof_gemm.py --which fc --axis time --reps 3 2>&1 | tail -1 >> $LOG
cat $LOG
