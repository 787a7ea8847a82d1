#!/bin/bash
# r02 call 52 (1 GPU): run-merged bulk loads in norm_cast (time axis); --set full captures of the grouped BandSplit GEMM, the
# band-axis Linear+skip, norm_cast (both axes) and the mask decoder's Conv1d+GLU GEMM.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensorcore_vs_oracle or graph_replay or fused_and_unfused" > gpurun_out/r02c52_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02c52_pytest.log
KREG='regex:^(lstm_|gemm_|norm_cast|istft|stft|band_|gn_|void lstm|void gemm|void stft|void norm)'
B="python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline --no-fp32"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" --csv --log-file gpurun_out/r02c52_ncu_launches_bench.csv $B > gpurun_out/r02c52_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
# first forward of the process: launch order stft, gn_finalize, band_norm_cast, gemm<8> (grouped band split), [gn_finalize, norm_cast, lstm, gemm<8>] x 12 ...
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:gemm_tc_kernel<\(int\)8|gemm_tc_kernel<8' -c 3 -o gpurun_out/r02c52_gemm8 -f $B > gpurun_out/r02c52_ncu_gemm8.log 2>&1; echo "ncu gemm8 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:norm_cast' -c 2 -o gpurun_out/r02c52_norm -f $B > gpurun_out/r02c52_ncu_norm.log 2>&1; echo "ncu norm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:gemm_tc_kernel<\(int\)3|gemm_tc_kernel<3' -s 20 -c 1 -o gpurun_out/r02c52_glu -f $B > gpurun_out/r02c52_ncu_glu.log 2>&1; echo "ncu glu rc=$?"
ls -la gpurun_out/r02c52*
