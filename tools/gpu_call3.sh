#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 -k "tensorcore" -s > gpurun_out/pytest_tc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tc.log
tail -n 15 gpurun_out/pytest_tc.log
timeout 900 python bench.py --precision fp16 --batch 64 --seconds 10 --steps 3 --warmup 3 --cpu-seconds 4 > gpurun_out/bench_tc_full.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_tc_full.log
tail -n 5 gpurun_out/bench_tc_full.log
