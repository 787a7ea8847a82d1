#!/bin/bash
# r02 call 8 (1 GPU): launch list of the bench command restricted to this library's kernels (one full forward), and
# --set full of the training step kernels (forward step, BPTT step, split-K weight-gradient GEMM, loss kernel).
mkdir -p gpurun_out
KREG='regex:^(lstm_|gemm_|norm_cast|istft|stft|band_stats|gn_|twiddle)'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 207 -c 207 --csv --log-file gpurun_out/r02c08_ncu_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c08_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
grep -c lstm_tc gpurun_out/r02c08_ncu_launches_bench.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4000 -c 6 -o gpurun_out/r02c08_train_steps -f \
  python tools/prof_train.py > gpurun_out/r02c08_ncu_train.log 2>&1; echo "ncu train rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"mrl1|kb8_transpose" -s 8 -c 3 -o gpurun_out/r02c08_train_misc -f \
  python tools/prof_train.py > gpurun_out/r02c08_ncu_train2.log 2>&1; echo "ncu train2 rc=$?"
ls -la gpurun_out/r02c08*.ncu-rep
