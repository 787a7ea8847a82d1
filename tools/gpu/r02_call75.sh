#!/bin/bash
# r02 call 75 (1 GPU): BPTT tile width chosen per block (32 where one wave of CTAs fits, else 64) vs 32 everywhere; training tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q -x > gpurun_out/r02c75_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r02c75_pytest.log
for BN in 0 32; do
BSRNN_BWD_BN=$BN timeout 600 python bench.py --config 5 --no-cpu-baseline --no-library-baseline 2>/dev/null | tail -1 > gpurun_out/r02c75_bench_cfg5_bn$BN.json; python -c "
import json; d=json.loads(open('gpurun_out/r02c75_bench_cfg5_bn$BN.json').read()); print('bn $BN', round(d['ms_per_step'],2), round(d['value'],1), d['loss'])"
done
