#!/bin/bash
# r02 call 51 (1 GPU): grouped BandSplit GEMM, shared normalised operand of the two mask-decoder MLP families, evict_last L2
# policy on the decoder's hidden activation: parity tests, then launch lists (kernels alone) with A/B switches.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q -x -s -k "band_split or decoder_statistics or tensorcore or graph_replay or verbatim or reference" > gpurun_out/r02c51_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "rel_l2|vs f32|statistics|passed|failed|Error" gpurun_out/r02c51_pytest.log | tail -24
KREG='regex:^(lstm_|gemm_|norm_cast|istft|stft|band_|gn_|void lstm|void gemm|void stft|void norm)'
B="python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline --no-fp32"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" --csv --log-file gpurun_out/r02c51_ncu_launches_bench.csv $B > gpurun_out/r02c51_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
BSRNN_HIDDEN_L2=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" --csv --log-file gpurun_out/r02c51_ncu_launches_nokeep.csv $B > /dev/null 2>&1; echo "launch list (no L2 keep) rc=$?"
BSRNN_BAND_SPLIT_GROUPED=0 BSRNN_MASKDEC_SHARED_NORM=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" --csv --log-file gpurun_out/r02c51_ncu_launches_ungrouped.csv $B > /dev/null 2>&1; echo "launch list (ungrouped, per-family norm) rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c51_bench_cfg2.json 2> gpurun_out/r02c51_bench_cfg2.err; echo "bench rc=$?"; cut -c 1-400 gpurun_out/r02c51_bench_cfg2.json
