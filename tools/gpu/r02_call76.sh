#!/bin/bash
# r02 call 76 (2 GPUs): DDP training step (config 5) of the final tree + the 2-rank config-2 line
mkdir -p gpurun_out
P=29741
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --config 5 > gpurun_out/r02c76_bench_cfg5_n2.json 2> gpurun_out/r02c76_bench_cfg5_n2.err; echo "cfg5 n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c76_bench_cfg5_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['split_ms'], d['loss'], d['impl_notes']['blstm'][:40])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+2)) bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c76_bench_n2.json 2> gpurun_out/r02c76_bench_n2.err; echo "n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c76_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','scaling','n_gpus')}, d['e2e']['value'], d.get('strong',{}).get('value'))
PY
