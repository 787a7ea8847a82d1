#!/bin/bash
# GPU call 40: input projection with bulk-store (shared -> global) epilogue: unit check, timing, parity, bench.
mkdir -p gpurun_out
timeout 120 python tools/gpu_check_tc.py > gpurun_out/call40_check_tc.log 2>&1; echo "check_tc rc=$?"; grep -v "lstm" gpurun_out/call40_check_tc.log | tail -8
G=gpurun_out/call40_gemm.log; : > $G
for ax in time freq; do
  timeout 120 python tools/prof_gemm.py --which inproj --axis $ax --reps 3 --nobias >> $G 2>&1
done
BSRNN_GEMM_DEBUG=2 timeout 120 python tools/prof_gemm.py --which inproj --axis time --reps 3 --nobias 2>&1 | sed "s/^/[debug=2] /" >> $G
BSRNN_GEMM_DEBUG=1 timeout 120 python tools/prof_gemm.py --which inproj --axis time --reps 3 --nobias 2>&1 | sed "s/^/[debug=1] /" >> $G
cat $G
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/call40_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/call40_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/call40_bench.json 2> gpurun_out/call40_bench.err; echo "bench rc=$?"; cut -c1-1000 gpurun_out/call40_bench.json; tail -3 gpurun_out/call40_bench.err
