#!/bin/bash
# GPU call 33 (1 GPU): bulk-copy latency/throughput micro-benchmark; StreamedEnhancer test; bench with pipelined e2e.
mkdir -p gpurun_out
timeout 120 tools/_bin/tma_bench > gpurun_out/call33_tma.log 2>&1; cat gpurun_out/call33_tma.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/call33_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/call33_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/call33_bench.json 2> gpurun_out/call33_bench.err; echo "bench rc=$?"; cut -c1-900 gpurun_out/call33_bench.json; tail -3 gpurun_out/call33_bench.err
