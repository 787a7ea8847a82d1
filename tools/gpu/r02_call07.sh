#!/bin/bash
# r02 call 7 (1 GPU): ncu evidence for round 2: launch list of the bench command (host launches, so every kernel is listed),
# --set full of the flag-group recurrence on both axes at config-2 size, and of the training step kernels.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 414 -c 420 --csv --log-file gpurun_out/r02c07_ncu_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c07_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
tail -2 gpurun_out/r02c07_ncu_launches_bench.csv | cut -c 1-200
for ax in time freq; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_flag -s 1 -c 1 -o gpurun_out/r02c07_lstm_flag_$ax -f \
    python tools/prof_lstm.py --B 64 --T 1001 --K 34 --axis $ax --flag --reps 2 > gpurun_out/r02c07_ncu_lstm_$ax.log 2>&1; echo "ncu lstm $ax rc=$?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel<6|gemm_tc_kernel<5" -s 2000 -c 4 -o gpurun_out/r02c07_train_steps -f \
  python tools/prof_train.py > gpurun_out/r02c07_ncu_train.log 2>&1; echo "ncu train rc=$?"
ls -la gpurun_out/*.ncu-rep
