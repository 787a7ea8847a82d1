#!/bin/bash
# r02 call 5 (1 GPU): tensor-core training with the step loops in C; config 5 bench; FlowSE training step timing.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q -s -x -k "tensorcore" > gpurun_out/r02c05_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "BLSTM block|tensor-core train|passed|failed|Error|error|assert" gpurun_out/r02c05_pytest.log | tail -12
for prec in fp16; do
  timeout 600 python bench.py --config 5 --train-precision $prec > gpurun_out/r02c05_bench_cfg5_$prec.json 2> gpurun_out/r02c05_bench_cfg5_$prec.err; echo "cfg5 $prec rc=$?"; tail -2 gpurun_out/r02c05_bench_cfg5_$prec.err
  python -c "
import json
d=json.loads(open('gpurun_out/r02c05_bench_cfg5_$prec.json').read().strip().splitlines()[-1])
print('$prec', d['ms_per_step'], d['value'], d['split_ms'], d['loss'], d['gpu_launches']/d['steps'])"
done
timeout 300 python tools/prof_train.py 2>&1 | grep -v "^---" | cut -c 1-190 | tail -50
