#!/bin/bash
# GPU call 27 (1 GPU): specialised straight-line input-projection epilogue + norm_cast with 8 loads in flight:
# parity (pytest -m gpu), GEMM timing, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/call27_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/call27_pytest_gpu.log
G=gpurun_out/call27_gemm.log; : > $G
for ax in time freq; do
  timeout 120 python tools/prof_gemm.py --which inproj --axis $ax --reps 3 --nobias >> $G 2>&1
done
cat $G
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/call27_bench.json 2> gpurun_out/call27_bench.err; echo "bench rc=$?"; cat gpurun_out/call27_bench.json; tail -3 gpurun_out/call27_bench.err
