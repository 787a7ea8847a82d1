#!/bin/bash
# r02 call 62 (8 GPUs): the contract's 8-rank launch: config 2 weak + strong partition in one line; config 5 DDP step.
mkdir -p gpurun_out
P=29731
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c62_bench_n8.json 2> gpurun_out/r02c62_bench_n8.err; echo "n8 rc=$?"; tail -2 gpurun_out/r02c62_bench_n8.err | cut -c 1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c62_bench_n8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','scaling','n_gpus')}, d['e2e'], d.get('strong'), d['per_rank'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((P+2)) bench.py --gpus 8 --config 5 > gpurun_out/r02c62_bench_cfg5_n8.json 2> gpurun_out/r02c62_bench_cfg5_n8.err; echo "cfg5 n8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c62_bench_cfg5_n8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['split_ms'], d['loss'])
PY
