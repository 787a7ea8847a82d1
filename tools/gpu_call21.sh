#!/bin/bash
# GPU call 21: GEMM pipeline depth 8 + bulk L2 prefetch: A/B against depth 4 at config-2 shapes, unit checks, ncu of the
# input-projection GEMM, bench.
mkdir -p gpurun_out
LOG=gpurun_out/call21_gemm.log
: > $LOG
for w in inproj fc; do for ax in time freq; do
  BSRNN_GEMM_STAGES=4 timeout 120 python tools/prof_gemm.py --which $w --axis $ax --reps 3 >> $LOG 2>&1
  timeout 120 python tools/prof_gemm.py --which $w --axis $ax --reps 3 >> $LOG 2>&1
done; done
timeout 600 python tools/gpu_check_tc.py 2>&1 | grep -E "gemm|rror" >> $LOG
cat $LOG | tail -40
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/call21_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/call21_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/call21_bench.json 2> gpurun_out/call21_bench.err; echo "bench rc=$?"; cat gpurun_out/call21_bench.json; tail -3 gpurun_out/call21_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -o gpurun_out/call21_gemm_inproj_full python tools/prof_gemm.py --which inproj --axis time --reps 1 > gpurun_out/call21_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -o gpurun_out/call21_gemm_fc_full python tools/prof_gemm.py --which fc --axis time --reps 1 >> gpurun_out/call21_ncu_gemm.log 2>&1; echo "ncu gemm fc rc=$?"
