#!/bin/bash
# GPU call 34 (1 GPU): division-free tile iterator in the GEMM control warps; graph-cache eviction fix; parity + timings.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/call34_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/call34_pytest_gpu.log
G=gpurun_out/call34_gemm.log; : > $G
for ax in time freq; do
  timeout 120 python tools/prof_gemm.py --which inproj --axis $ax --reps 3 --nobias >> $G 2>&1
  timeout 120 python tools/prof_gemm.py --which fc --axis $ax --reps 3 >> $G 2>&1
done
cat $G
timeout 120 python tools/gpu_check_tc.py > gpurun_out/call34_check_tc.log 2>&1; echo "check_tc rc=$?"; tail -8 gpurun_out/call34_check_tc.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/call34_bench.json 2> gpurun_out/call34_bench.err; echo "bench rc=$?"; cut -c1-1000 gpurun_out/call34_bench.json; tail -3 gpurun_out/call34_bench.err
