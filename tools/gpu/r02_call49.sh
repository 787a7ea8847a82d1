#!/bin/bash
# r02 call 49 (1 GPU): tensor-core BandSplit + decoder statistics from the last Linear epilogue: unit tests, config-2 bench,
# then --set full captures of the small kernels of one step (STFT, band split, statistics, iSTFT, mask-decoder GEMMs, norm_cast).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "band_split_tc or decoder_statistics or tensorcore_vs_oracle or graph_replay" -s > gpurun_out/r02c49_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "rel_l2|vs f32|statistics|passed|failed|Error" gpurun_out/r02c49_pytest.log | tail -15
timeout 900 python bench.py --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c49_bench_cfg2.json 2> gpurun_out/r02c49_bench_cfg2.err; echo "bench rc=$?"
BSRNN_BAND_SPLIT_TC=0 timeout 900 python bench.py --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c49_bench_cfg2_bs_f32.json 2> gpurun_out/r02c49_bench_cfg2_bs_f32.err; echo "bench (f32 band split) rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r02c49_bench_cfg2.json','gpurun_out/r02c49_bench_cfg2_bs_f32.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'], {k:round(v,2) for k,v in d['roofline']['regions_ms_per_step'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
B="python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline --no-fp32"
KREG='regex:^(lstm_|gemm_|norm_cast|istft|stft|band_|gn_|void lstm|void gemm|void stft|void norm)'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" --csv --log-file gpurun_out/r02c49_ncu_launches_bench.csv \
  $B > gpurun_out/r02c49_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^(band_|istft|stft|gn_stats)' -c 7 -o gpurun_out/r02c49_small -f \
  $B > gpurun_out/r02c49_ncu_small.log 2>&1; echo "ncu small rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:gemm_tc_kernel<\(int\)2|gemm_tc_kernel<\(int\)3|gemm_tc_kernel<2|gemm_tc_kernel<3' -s 40 -c 4 -o gpurun_out/r02c49_maskdec -f \
  $B > gpurun_out/r02c49_ncu_maskdec.log 2>&1; echo "ncu maskdec rc=$?"
ls -la gpurun_out/r02c49*
tail -3 gpurun_out/r02c49_ncu_maskdec.log
