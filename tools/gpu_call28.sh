#!/bin/bash
# GPU call 28 (1 GPU): MUFU and TMEM-read micro-benchmarks; input projection with staggered epilogue groups; parity.
mkdir -p gpurun_out
timeout 60 tools/_bin/mufu_bench > gpurun_out/call28_mufu.log 2>&1; cat gpurun_out/call28_mufu.log
timeout 60 tools/_bin/tmem_bench > gpurun_out/call28_tmem.log 2>&1; cat gpurun_out/call28_tmem.log
G=gpurun_out/call28_gemm.log; : > $G
for ax in time freq; do
  timeout 120 python tools/prof_gemm.py --which inproj --axis $ax --reps 3 --nobias >> $G 2>&1
done
cat $G
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/call28_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/call28_pytest_gpu.log
