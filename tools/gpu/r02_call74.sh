#!/bin/bash
# r02 call 74 (1 GPU): BPTT output tile width at the training batch (more, smaller CTAs per step): gradient test + config 5 A/B
mkdir -p gpurun_out
for BN in 32 64 128; do
BSRNN_BWD_BN=$BN timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q -x -k "train_step_tensorcore_gradients or blstm_block_tensorcore" 2>&1 | tail -1
BSRNN_BWD_BN=$BN timeout 600 python bench.py --config 5 --no-cpu-baseline --no-library-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bn $BN', round(d['ms_per_step'],2), round(d['value'],1), d['loss'])"
done
