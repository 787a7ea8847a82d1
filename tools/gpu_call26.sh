#!/bin/bash
# GPU call 26 (1 GPU): `ncu --set full --import-source on` of the input-projection GEMM (bias folded) and of norm_cast_kb8
# at BASELINE config 2 shapes.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 1 -o gpurun_out/call26_gemm_inproj \
  python tools/prof_gemm.py --which inproj --axis time --reps 1 --nobias > gpurun_out/call26_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"; tail -2 gpurun_out/call26_ncu_gemm.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:norm_cast_kb8 -c 1 -o gpurun_out/call26_norm_cast \
  python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/call26_ncu_norm.log 2>&1; echo "ncu norm rc=$?"; tail -2 gpurun_out/call26_ncu_norm.log
ls -la gpurun_out/
