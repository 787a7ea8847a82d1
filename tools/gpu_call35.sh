#!/bin/bash
# GPU call 35 (1 GPU): control roles under a single elect_one region (GEMM + recurrence), unrolled MMA loop of the
# input projection: parity + timings.
mkdir -p gpurun_out
G=gpurun_out/call35_gemm.log; : > $G
for ax in time freq; do
  timeout 120 python tools/prof_gemm.py --which inproj --axis $ax --reps 3 --nobias >> $G 2>&1
  timeout 120 python tools/prof_gemm.py --which fc --axis $ax --reps 3 >> $G 2>&1
done
cat $G
LOG=gpurun_out/call35_lstm.log; : > $LOG
for ax in time freq; do
  timeout 120 python tools/prof_lstm.py --B 64 --T 1001 --K 34 --axis $ax --slots 3 --reps 3 --trace >> $LOG 2>&1
done
grep -E "CHECK|FAILED|ms,|producer|mma  |epilogue|rror" $LOG
timeout 120 python tools/gpu_check_tc.py > gpurun_out/call35_check_tc.log 2>&1; echo "check_tc rc=$?"; grep -c OK gpurun_out/call35_check_tc.log; grep -i "fail\|error" gpurun_out/call35_check_tc.log | head
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/call35_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/call35_pytest_gpu.log
