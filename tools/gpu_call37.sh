#!/bin/bash
# GPU call 37: where does the input projection's time go?  A/B: normal / stores kept in L2 / no stores / no L2 prefetch.
mkdir -p gpurun_out
G=gpurun_out/call37_gemm.log; : > $G
for d in 0 1 2; do
  BSRNN_GEMM_DEBUG=$d timeout 120 python tools/prof_gemm.py --which inproj --axis time --reps 3 --nobias 2>&1 | sed "s/^/[debug=$d] /" >> $G
done
BSRNN_GEMM_PFDIST=0 timeout 120 python tools/prof_gemm.py --which inproj --axis time --reps 3 --nobias 2>&1 | sed "s/^/[pfdist=0] /" >> $G
BSRNN_GEMM_PFDIST=1 timeout 120 python tools/prof_gemm.py --which inproj --axis time --reps 3 --nobias 2>&1 | sed "s/^/[pfdist=1] /" >> $G
BSRNN_GEMM_PFDIST=6 timeout 120 python tools/prof_gemm.py --which inproj --axis time --reps 3 --nobias 2>&1 | sed "s/^/[pfdist=6] /" >> $G
BSRNN_GEMM_DEBUG=2 BSRNN_GEMM_PFDIST=0 timeout 120 python tools/prof_gemm.py --which inproj --axis time --reps 3 --nobias 2>&1 | sed "s/^/[debug=2 pfdist=0] /" >> $G
cat $G
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/call37_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/call37_pytest_gpu.log
