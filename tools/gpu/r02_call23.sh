#!/bin/bash
# r02 call 23 (1 GPU): ncu evidence for the fused schedule: launch list of one bench step (host launches) and --set full
# captures of one time-axis + one band-axis block (norm_cast, lstm_fused, Linear+skip GEMM).
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 195 -c 200 --csv --log-file gpurun_out/r02c23_ncu_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c23_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"lstm_fused|gemm_tc_kernel<1|norm_cast_kb8" -s 12 -c 6 -o gpurun_out/r02c23_block -f \
  python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c23_ncu_block.log 2>&1; echo "ncu block rc=$?"
ls -la gpurun_out/r02c23*
