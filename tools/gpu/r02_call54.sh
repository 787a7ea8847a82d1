#!/bin/bash
# r02 call 54 (1 GPU): norm_cast block size (default now half blocks: 64 rows at N = 196, 32 at N = 384), mask decoder on half
# of the SMs per stream: launch lists with A/B switches, FlowSE config 4 bench, full gpu test suite.
mkdir -p gpurun_out
KREG='regex:^(lstm_|gemm_|norm_cast|istft|stft|band_|gn_|void lstm|void gemm|void stft|void norm)'
B="python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline --no-fp32"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" --csv --log-file gpurun_out/r02c54_ncu_launches_bench.csv $B > gpurun_out/r02c54_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
BSRNN_PACK_ROWS=32 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" --csv --log-file gpurun_out/r02c54_ncu_launches_rows32.csv $B > /dev/null 2>&1; echo "launch list (32-row blocks) rc=$?"
for S in 0 74; do
BSRNN_MASKDEC_SMS=$S timeout 600 python bench.py --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c54_bench_cfg2_sms$S.json 2> gpurun_out/r02c54_bench_cfg2_sms$S.err; echo "bench sms=$S rc=$?"
done
timeout 900 python bench.py --config 4 --no-cpu-baseline --no-library-baseline > gpurun_out/r02c54_bench_cfg4.json 2> gpurun_out/r02c54_bench_cfg4.err; echo "cfg4 rc=$?"
BSRNN_PACK_ROWS=64 timeout 900 python bench.py --config 4 --no-cpu-baseline --no-library-baseline > gpurun_out/r02c54_bench_cfg4_rows64.json 2> /dev/null; echo "cfg4 rows64 rc=$?"
python - <<'PY'
import json
for f in ('cfg2_sms0','cfg2_sms74','cfg4','cfg4_rows64'):
    try:
        d=json.loads(open(f'gpurun_out/r02c54_bench_{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],2), round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], {k:round(v,2) for k,v in d.get('roofline',{}).get('regions_ms_per_step',{}).items()})
    except Exception as e: print(f, 'ERR', e)
PY
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02c54_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02c54_pytest.log
