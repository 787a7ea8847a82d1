#!/bin/bash
# GPU call 16: residual epilogue with batched reads + next-tile prefetch: unit checks, gpu tests, bench (graph / host).
mkdir -p gpurun_out
timeout 600 python tools/gpu_check_tc.py 2>&1 | grep -E "gemm resid|gemm rows   M=1000|rror" | tee gpurun_out/call16_unit.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/call16_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/call16_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/call16_bench.json 2> gpurun_out/call16_bench.err; echo "bench rc=$?"; cat gpurun_out/call16_bench.json; tail -3 gpurun_out/call16_bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/call16_bench_nograph.json 2>> gpurun_out/call16_bench.err; cat gpurun_out/call16_bench_nograph.json
