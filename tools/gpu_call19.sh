#!/bin/bash
# GPU call 19: evidence set on the restored tree: all gpu tests (inference + training), bench, ncu launch list of the bench
# command, ncu --set full of the recurrence and GEMM kernels, train-step timing.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/call19_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/call19_pytest_gpu.log; tail -3 gpurun_out/call19_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/call19_bench.json 2> gpurun_out/call19_bench.err; echo "bench rc=$?"; cat gpurun_out/call19_bench.json; tail -3 gpurun_out/call19_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/call19_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/call19_ncu_bench.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/call19_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tc -s 1 -c 2 -o gpurun_out/call19_lstm_full python tools/prof_lstm.py --B 64 --T 1001 --K 34 --axis time --reps 2 > gpurun_out/call19_ncu_lstm.log 2>&1; echo "ncu lstm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 30 -c 4 -o gpurun_out/call19_gemm_full python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-graph --batch 16 > gpurun_out/call19_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 python tools/bench_train.py > gpurun_out/call19_train_1gpu.json 2> gpurun_out/call19_train.err; cat gpurun_out/call19_train_1gpu.json; tail -3 gpurun_out/call19_train.err
ls -la gpurun_out | tail -12
