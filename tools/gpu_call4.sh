#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_check_tc.py > gpurun_out/check_tc.log 2>&1; echo "rc=$?" >> gpurun_out/check_tc.log
tail -n 18 gpurun_out/check_tc.log
timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 -k "tensorcore" -s > gpurun_out/pytest_tc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tc.log
tail -n 6 gpurun_out/pytest_tc.log
timeout 900 python bench.py --precision fp16 --batch 64 --seconds 10 --steps 3 --warmup 3 --cpu-seconds 4 > gpurun_out/bench_tc_full.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_tc_full.log
tail -n 3 gpurun_out/bench_tc_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --precision fp16 --batch 16 --seconds 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?" >> gpurun_out/ncu_bench.log
tail -n 3 gpurun_out/ncu_bench.log
