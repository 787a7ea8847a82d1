#!/bin/bash
# r02 call 72 (1 GPU): band-innermost M walk of the band-axis Linear+skip without a tensor map (FlowSE, two N tiles): bit check,
# FlowSE tests, config 4 A/B
mkdir -p gpurun_out
python tools/check_fc_tma.py 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_gpu_training.py -m gpu -q -x -k "flowse or FlowSE or blstm_block or tensorcore" > gpurun_out/r02c72_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02c72_pytest.log
for V in 1 0; do
BSRNN_FC_BAND_INNER=$V timeout 900 python bench.py --config 4 --no-cpu-baseline --no-library-baseline > gpurun_out/r02c72_bench_cfg4_inner$V.json 2> gpurun_out/r02c72_bench_cfg4_inner$V.err; echo "cfg4 inner=$V rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02c72_bench_cfg4_inner$V.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],1), round(d['value'],2), d['clocks']['sm_mhz'], d.get('regions_ms_per_eval') or d.get('roofline',{}).get('regions_ms_per_step'))"
done
