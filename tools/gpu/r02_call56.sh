#!/bin/bash
# r02 call 56 (1 GPU): after the fix of the bulk-store tanh epilogue (26th core of each hidden tile): full gpu suite, launch list, bench
mkdir -p gpurun_out
timeout 300 python tools/debug_fullsize.py 10 64 2>&1 | grep -E "switches|Error" | cut -c 1-200
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02c56_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02c56_pytest.log
KREG='regex:^(lstm_|gemm_|norm_cast|istft|stft|band_|gn_|void lstm|void gemm|void stft|void norm)'
B="python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline --no-fp32"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" --csv --log-file gpurun_out/r02c56_ncu_launches_bench.csv $B > gpurun_out/r02c56_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c56_bench_cfg2.json 2> gpurun_out/r02c56_bench_cfg2.err; echo "bench rc=$?"; cut -c 1-330 gpurun_out/r02c56_bench_cfg2.json
