#!/bin/bash
# r02 call 69 (1 GPU): band-fastest grids of the BandSplit kernels (operand builder, f32 kernel): tests, launch list, FlowSE config 4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q -x > gpurun_out/r02c69_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02c69_pytest.log
KREG='regex:^(lstm_|gemm_|norm_cast|istft|stft|band_|gn_|void lstm|void gemm|void stft|void norm)'
B="python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline --no-fp32"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" --csv --log-file gpurun_out/r02c69_ncu_launches_bench.csv $B > gpurun_out/r02c69_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 python bench.py --config 4 --no-cpu-baseline --no-library-baseline > gpurun_out/r02c69_bench_cfg4.json 2> gpurun_out/r02c69_bench_cfg4.err; echo "cfg4 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02c69_bench_cfg4.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],1), round(d['value'],2), d['clocks']['sm_mhz'])"
