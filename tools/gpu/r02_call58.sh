#!/bin/bash
# r02 call 58 (1 GPU): training forward on the fused layer kernel (activated gates + c_t saved by its epilogue): tests, config 5 A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q -x -s > gpurun_out/r02c58_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "BLSTM block TC|tensor-core train step|passed|failed|Error" gpurun_out/r02c58_pytest.log | tail -14
for V in 1 0; do
BSRNN_TRAIN_FUSED=$V timeout 600 python bench.py --config 5 --no-cpu-baseline --no-library-baseline > gpurun_out/r02c58_bench_cfg5_fused$V.json 2> gpurun_out/r02c58_bench_cfg5_fused$V.err; echo "cfg5 fused=$V rc=$?"
done
python - <<'PY'
import json
for f in ('fused1','fused0'):
    try:
        d=json.loads(open(f'gpurun_out/r02c58_bench_cfg5_{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],2), round(d['value'],1), d.get('split_ms'), d['gpu_launches'], d['loss'])
    except Exception as e: print(f, 'ERR', e); print(open(f'gpurun_out/r02c58_bench_cfg5_{f}.err').read()[-1500:])
PY
