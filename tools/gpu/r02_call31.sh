#!/bin/bash
# r02 call 31: launch list of FlowSE network evaluations at config-4 size (32x10s), our kernels only.
mkdir -p gpurun_out
KREG='regex:^(lstm_|gemm_|norm_cast|istft|stft|band_|gn_|conv5x5|euler|axpy|complex_mask|time_embed|void gemm_|void stft)'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 400 --csv --log-file gpurun_out/r02c31_ncu_launches_flowse.csv \
  python tools/bench_flowse.py --batch 32 --nfe 1 --reps 1 > gpurun_out/r02c31_flowse_under_ncu.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/r02c31_flowse_under_ncu.log
