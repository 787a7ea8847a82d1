#!/bin/bash
# r02 call 64 (1 GPU): STFT / iSTFT staging loops with all frames' loads in flight: unit tests + launch list + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_training.py -m gpu -q -x -k "stft or istft or fullsize or multires or tensorcore_vs_oracle or flowse_vs_golden" > gpurun_out/r02c64_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02c64_pytest.log
KREG='regex:^(lstm_|gemm_|norm_cast|istft|stft|band_|gn_|void lstm|void gemm|void stft|void norm)'
B="python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline --no-fp32"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" --csv --log-file gpurun_out/r02c64_ncu_launches_bench.csv $B > gpurun_out/r02c64_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
