#!/bin/bash
# r02 call 57 (1 GPU): training step with the batched per-band ops (BandSplit / MaskDecoder as batched GEMMs): tests + config 5 A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q -x > gpurun_out/r02c57_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02c57_pytest.log
for V in 1 0; do
BSRNN_TRAIN_BATCHED_BANDS=$V timeout 600 python bench.py --config 5 --no-cpu-baseline --no-library-baseline > gpurun_out/r02c57_bench_cfg5_batched$V.json 2> gpurun_out/r02c57_bench_cfg5_batched$V.err; echo "cfg5 batched=$V rc=$?"
done
python - <<'PY'
import json
for f in ('batched1','batched0'):
    try:
        d=json.loads(open(f'gpurun_out/r02c57_bench_cfg5_{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],2), round(d['value'],1), d.get('split_ms'), d['gpu_launches'], d['loss'])
    except Exception as e: print(f, 'ERR', e)
PY
