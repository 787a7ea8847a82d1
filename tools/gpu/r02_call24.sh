#!/bin/bash
# r02 call 24 (1 GPU): launch list of one bench step of the fused schedule (our kernels only), --set full of the Linear+skip GEMM.
mkdir -p gpurun_out
KREG='regex:^(lstm_|gemm_|norm_cast|istft|stft|band_|gn_)'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 195 -c 195 --csv --log-file gpurun_out/r02c24_ncu_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c24_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 2 -o gpurun_out/r02c24_fc -f \
  python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c24_ncu_fc.log 2>&1; echo "ncu fc rc=$?"
ls -la gpurun_out/r02c24*
