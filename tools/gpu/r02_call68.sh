#!/bin/bash
# r02 call 68 (1 GPU): balanced column partition of the Linear+skip / BandSplit epilogue warps: bit check, parity, launch list
mkdir -p gpurun_out
python tools/check_fc_tma.py 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r02c68_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02c68_pytest.log
KREG='regex:^(lstm_|gemm_|norm_cast|istft|stft|band_|gn_|void lstm|void gemm|void stft|void norm)'
B="python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline --no-fp32"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" --csv --log-file gpurun_out/r02c68_ncu_launches_bench.csv $B > gpurun_out/r02c68_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
