#!/bin/bash
# first GPU call: stage check, gpu tests, small bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
timeout 600 python tools/gpu_check.py > gpurun_out/check.log 2>&1; echo "check rc=$?" >> gpurun_out/check.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --batch 8 --seconds 4 --steps 2 --warmup 1 --cpu-seconds 2 > gpurun_out/bench_small.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_small.log
tail -5 gpurun_out/check.log gpurun_out/pytest_gpu.log gpurun_out/bench_small.log
