#!/bin/bash
# r02 call 53 (1 GPU): 2-D tensor-map copies (band-axis Linear+skip, grouped BandSplit stores, band-axis norm_cast loads):
# parity tests, memcheck of a small forward, launch lists with A/B switches.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q -x -s > gpurun_out/r02c53_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "verbatim|vs f32|statistics|passed|failed|Error" gpurun_out/r02c53_pytest.log | tail -24
python tools/check_fc_tma.py 2>&1 | tail -5
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/r02c53_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02c53_memcheck.log
KREG='regex:^(lstm_|gemm_|norm_cast|istft|stft|band_|gn_|void lstm|void gemm|void stft|void norm)'
B="python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline --no-fp32"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" --csv --log-file gpurun_out/r02c53_ncu_launches_bench.csv $B > gpurun_out/r02c53_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
BSRNN_FC_TMAP=0 BSRNN_PACK_TMAP=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" --csv --log-file gpurun_out/r02c53_ncu_launches_notmap.csv $B > /dev/null 2>&1; echo "launch list (no tensor maps) rc=$?"
BSRNN_PACK_ROWS=64 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" --csv --log-file gpurun_out/r02c53_ncu_launches_rows64.csv $B > /dev/null 2>&1; echo "launch list (64-row norm_cast blocks) rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c53_bench_cfg2.json 2> gpurun_out/r02c53_bench_cfg2.err; echo "bench rc=$?"; cut -c 1-330 gpurun_out/r02c53_bench_cfg2.json
