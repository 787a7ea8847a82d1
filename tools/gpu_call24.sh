#!/bin/bash
# GPU call 24 (1 GPU): write-bandwidth micro-benchmark; recurrence v7 (CTA pairs, cta_group::2) parity + timing vs v5;
# Linear+skip GEMM L2-prefetch distance A/B; bench with both recurrence schedules.
mkdir -p gpurun_out
timeout 120 tools/_bin/wbench > gpurun_out/call24_wbench.log 2>&1; echo "wbench rc=$?"; cat gpurun_out/call24_wbench.log
LOG=gpurun_out/call24_lstm_v7.log
: > $LOG
P="timeout 120 python tools/prof_lstm.py --ver 7"
$P --B 12 --T 40 --K 34 --axis time --slots 1 --check --reps 1 >> $LOG 2>&1 || echo "FAILED time slots=1 rc=$?" >> $LOG
if grep -q "OK" $LOG; then
  $P --B 10 --T 40 --K 34 --axis time --slots 2 --check --reps 1 >> $LOG 2>&1 || echo "FAILED time odd tiles slots=2 rc=$?" >> $LOG
  $P --B 3 --T 300 --K 34 --axis freq --slots 3 --check --reps 1 >> $LOG 2>&1 || echo "FAILED freq slots=3 rc=$?" >> $LOG
  $P --B 40 --T 60 --K 34 --axis time --slots 3 --maxcl 2 --check --reps 1 >> $LOG 2>&1 || echo "FAILED multi-group time" >> $LOG
  $P --B 40 --T 60 --K 34 --axis freq --slots 3 --maxcl 2 --check --reps 1 >> $LOG 2>&1 || echo "FAILED multi-group freq" >> $LOG
  for ax in time freq; do
    for sl in 2 3; do
      $P --B 64 --T 1001 --K 34 --axis $ax --slots $sl --reps 3 >> $LOG 2>&1
    done
    $P --B 64 --T 1001 --K 34 --axis $ax --slots 3 --reps 2 --trace >> $LOG 2>&1
    timeout 120 python tools/prof_lstm.py --ver 5 --B 64 --T 1001 --K 34 --axis $ax --slots 3 --reps 3 >> $LOG 2>&1
  done
fi
grep -E "CHECK|FAILED|ms,|co-resident|producer|mma  |epilogue|rror" $LOG | tail -40
G=gpurun_out/call24_gemm_fc.log; : > $G
for d in 3 1 0; do BSRNN_GEMM_PFDIST=$d timeout 120 python tools/prof_gemm.py --which fc --axis time --reps 3 2>&1 | sed "s/^/[pfdist=$d] /" >> $G; done
BSRNN_GEMM_PFDIST=1 timeout 120 python tools/prof_gemm.py --which fc --axis freq --reps 3 2>&1 | sed "s/^/[pfdist=1 freq] /" >> $G
BSRNN_GEMM_PFDIST=3 timeout 120 python tools/prof_gemm.py --which fc --axis freq --reps 3 2>&1 | sed "s/^/[pfdist=3 freq] /" >> $G
cat $G
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/call24_bench_v5.json 2> gpurun_out/call24_bench_v5.err; echo "bench v5 rc=$?"; cat gpurun_out/call24_bench_v5.json; tail -3 gpurun_out/call24_bench_v5.err
if grep -q "FAIL\|rror" $LOG; then echo "v7 not clean: skipping v7 bench"; else
  BSRNN_LSTM_VER=7 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/call24_bench_v7.json 2> gpurun_out/call24_bench_v7.err; echo "bench v7 rc=$?"; cat gpurun_out/call24_bench_v7.json; tail -3 gpurun_out/call24_bench_v7.err
  BSRNN_LSTM_VER=7 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/call24_pytest_gpu_v7.log 2>&1; echo "pytest v7 rc=$?"; tail -3 gpurun_out/call24_pytest_gpu_v7.log
fi
