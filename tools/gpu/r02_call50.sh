#!/bin/bash
# r02 call 50 (1 GPU): run-merged bulk copies in the Linear+skip epilogue, bulk-store tanh GEMM of the mask decoder, iSTFT block
# of 7 hops: parity tests, A/B benches, launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q -x -s > gpurun_out/r02c50_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "rel_l2|vs f32|statistics|passed|failed|Error" gpurun_out/r02c50_pytest.log | tail -30
python tools/check_fc_tma.py 2>&1 | tail -5
B="python bench.py --no-cpu-baseline --no-library-baseline --no-fp32"
timeout 600 $B > gpurun_out/r02c50_bench_cfg2.json 2> gpurun_out/r02c50_bench_cfg2.err; echo "bench rc=$?"
BSRNN_FC_RUNS=0 timeout 600 $B > gpurun_out/r02c50_bench_cfg2_noruns.json 2>/dev/null; echo "bench noruns rc=$?"
BSRNN_TANH_BULK=0 timeout 600 $B > gpurun_out/r02c50_bench_cfg2_notanhbulk.json 2>/dev/null; echo "bench notanhbulk rc=$?"
python - <<'PY'
import json
for f in ('cfg2','cfg2_noruns','cfg2_notanhbulk'):
    try:
        d=json.loads(open(f'gpurun_out/r02c50_bench_{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],2), round(d['value']), round(d['e2e']['value']), d['clocks']['sm_mhz'], {k:round(v,2) for k,v in d['roofline']['regions_ms_per_step'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
KREG='regex:^(lstm_|gemm_|norm_cast|istft|stft|band_|gn_|void lstm|void gemm|void stft|void norm)'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" --csv --log-file gpurun_out/r02c50_ncu_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c50_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
