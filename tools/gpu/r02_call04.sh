#!/bin/bash
# r02 call 4 (1 GPU): loss kernel, FlowSE training step, tensor-core BLSTM block fwd/bwd, config-5 bench in both precisions.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q -s -x -k "multires or flowse or tensorcore or inference_sees or adamw_v2 or trainer" > gpurun_out/r02c04_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "grad rel|forward_step|BLSTM block|tensor-core train|passed|failed|Error|error|assert" gpurun_out/r02c04_pytest.log | tail -30
for prec in fp16 fp32; do
  timeout 600 python bench.py --config 5 --train-precision $prec > gpurun_out/r02c04_bench_cfg5_$prec.json 2> gpurun_out/r02c04_bench_cfg5_$prec.err; echo "cfg5 $prec rc=$?"; tail -2 gpurun_out/r02c04_bench_cfg5_$prec.err
  python -c "
import json
d=json.loads(open('gpurun_out/r02c04_bench_cfg5_$prec.json').read().strip().splitlines()[-1])
print('$prec', d['ms_per_step'], d['value'], d['split_ms'], d['loss'], d['gpu_launches']/d['steps'])"
done
