#!/bin/bash
# r02 call 61 (2 GPUs): 2-rank launch of the committed tree (config 2 weak + strong partition, config 5 DDP step) and memcheck of
# the new training-forward kernels (fused layer kernel in save mode + transpose) on one small block.
mkdir -p gpurun_out
P=29721
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c61_bench_n2.json 2> gpurun_out/r02c61_bench_n2.err; echo "n2 rc=$?"; tail -2 gpurun_out/r02c61_bench_n2.err | cut -c 1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c61_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','scaling','n_gpus')}, d['e2e'], d.get('strong'), d['per_rank'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+2)) bench.py --gpus 2 --config 5 > gpurun_out/r02c61_bench_cfg5_n2.json 2> gpurun_out/r02c61_bench_cfg5_n2.err; echo "cfg5 n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c61_bench_cfg5_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['split_ms'], d['loss'])
PY
CUDA_VISIBLE_DEVICES=0 timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_training.py -m gpu -q -x -k "blstm_block_tensorcore and (1-23-34-196 or 8-101-6-196)" > gpurun_out/r02c61_memcheck_train_fused.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02c61_memcheck_train_fused.log | tail -3
