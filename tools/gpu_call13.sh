#!/bin/bash
# GPU call 13: 8-warp GEMM epilogues + coalesced residual epilogue, vectorised norm_cast, CUDA-graph forward.
mkdir -p gpurun_out
timeout 600 python tools/gpu_check_tc.py > gpurun_out/call13_unit_checks.log 2>&1; echo "unit rc=$?"; grep -E "gemm|lstm|rror" gpurun_out/call13_unit_checks.log | tail -30
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/call13_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/call13_pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/call13_bench.json 2> gpurun_out/call13_bench.err; echo "bench rc=$?"; cat gpurun_out/call13_bench.json; tail -5 gpurun_out/call13_bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/call13_bench_nograph.json 2>> gpurun_out/call13_bench.err; cat gpurun_out/call13_bench_nograph.json
