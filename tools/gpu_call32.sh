#!/bin/bash
# GPU call 32 (1 GPU): state of the tree with recurrence v8 default: pytest -m gpu, bench (with CPU baseline), reference
# arm, ncu launch list of the bench command, ncu --set full of the recurrence (time axis) and of the input projection.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/call32_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/call32_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/call32_bench.json 2> gpurun_out/call32_bench.err; echo "bench rc=$?"; cat gpurun_out/call32_bench.json; tail -3 gpurun_out/call32_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/call32_bench_ref.json 2> gpurun_out/call32_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/call32_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/call32_ncu_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/call32_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_kernel -c 1 -o gpurun_out/call32_lstm_v8_time \
  python tools/prof_lstm.py --B 64 --T 1001 --K 34 --axis time --slots 3 --reps 1 > gpurun_out/call32_ncu_lstm.log 2>&1; echo "ncu lstm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 1 -o gpurun_out/call32_gemm_inproj \
  python tools/prof_gemm.py --which inproj --axis time --reps 1 --nobias > gpurun_out/call32_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
ls -la gpurun_out | tail -12
