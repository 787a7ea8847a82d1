#!/bin/bash
# GPU call 29 (1 GPU): converged control warps (elect_one around MMA / bulk-copy issue) in the GEMM and the v5 recurrence:
# parity, kernel timings, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/call29_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/call29_pytest_gpu.log
G=gpurun_out/call29_gemm.log; : > $G
for ax in time freq; do
  timeout 120 python tools/prof_gemm.py --which inproj --axis $ax --reps 3 --nobias >> $G 2>&1
  timeout 120 python tools/prof_gemm.py --which fc --axis $ax --reps 3 >> $G 2>&1
done
cat $G
LOG=gpurun_out/call29_lstm.log; : > $LOG
for ax in time freq; do
  timeout 120 python tools/prof_lstm.py --ver 5 --B 64 --T 1001 --K 34 --axis $ax --slots 3 --reps 3 --trace >> $LOG 2>&1
done
timeout 120 python tools/prof_lstm.py --ver 5 --B 40 --T 60 --K 34 --axis time --slots 3 --maxcl 2 --check --reps 1 >> $LOG 2>&1
timeout 120 python tools/prof_lstm.py --ver 5 --B 3 --T 300 --K 34 --axis freq --slots 3 --check --reps 1 >> $LOG 2>&1
grep -E "CHECK|FAILED|ms,|producer|mma  |epilogue|rror" $LOG
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/call29_bench.json 2> gpurun_out/call29_bench.err; echo "bench rc=$?"; cat gpurun_out/call29_bench.json; tail -3 gpurun_out/call29_bench.err
