#!/bin/bash
# GPU call 25 (1 GPU): parity after folding the LSTM bias into the input-projection GEMM (no bias epilogue, pipelined
# TMEM loads), GEMM timings (inproj with/without bias epilogue, Linear+skip prefetch off), bench, and one
# `ncu --set full --import-source on` capture of the v5 recurrence on the band axis (source page -> where the epilogue stalls).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/call25_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/call25_pytest_gpu.log
G=gpurun_out/call25_gemm.log; : > $G
for ax in time freq; do
  timeout 120 python tools/prof_gemm.py --which inproj --axis $ax --reps 3 >> $G 2>&1
  timeout 120 python tools/prof_gemm.py --which inproj --axis $ax --reps 3 --nobias >> $G 2>&1
  timeout 120 python tools/prof_gemm.py --which fc --axis $ax --reps 3 >> $G 2>&1
  BSRNN_GEMM_PFDIST=1 timeout 120 python tools/prof_gemm.py --which fc --axis $ax --reps 3 2>&1 | sed "s/^/[pfdist=1] /" >> $G
done
cat $G
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/call25_bench.json 2> gpurun_out/call25_bench.err; echo "bench rc=$?"; cat gpurun_out/call25_bench.json; tail -3 gpurun_out/call25_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_kernel -c 1 -o gpurun_out/call25_lstm_v5_freq \
  python tools/prof_lstm.py --ver 5 --B 64 --T 1001 --K 34 --axis freq --slots 3 --reps 1 > gpurun_out/call25_ncu_lstm.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/call25_ncu_lstm.log
ls -la gpurun_out/
