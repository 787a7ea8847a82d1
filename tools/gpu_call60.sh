#!/bin/bash
# GPU call 60 (1 GPU): last check of the committed tree: pytest -m gpu (43 tests), smoke, default bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/call60_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/call60_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/call60_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/call60_smoke.log
timeout 600 python bench.py > gpurun_out/call60_bench_default.json 2> gpurun_out/call60_bench_default.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/call60_bench_default.json)"; cut -c1-900 gpurun_out/call60_bench_default.json
