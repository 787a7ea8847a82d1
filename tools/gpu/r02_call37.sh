#!/bin/bash
# r02 call 37 (2 GPUs): the contract's multi-rank launch with the fused schedule: config 2 weak + strong in one line, the
# reference arm under torchrun, config 4 and config 5 at N = 2.
mkdir -p gpurun_out
P=29611
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02c37_bench_n2.json 2> gpurun_out/r02c37_bench_n2.err; echo "n2 rc=$?"; tail -3 gpurun_out/r02c37_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c37_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','scaling','n_gpus')}, d['e2e'], d.get('strong'), d['per_rank'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | tail -1 | cut -c 1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+2)) bench.py --gpus 2 --config 5 > gpurun_out/r02c37_bench_cfg5_n2.json 2> gpurun_out/r02c37_bench_cfg5_n2.err; echo "cfg5 n2 rc=$?"; tail -2 gpurun_out/r02c37_bench_cfg5_n2.err | cut -c 1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c37_bench_cfg5_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['split_ms'], d['loss'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+3)) bench.py --gpus 2 --config 4 --steps 1 --warmup 1 --no-cpu-baseline --no-library-baseline > gpurun_out/r02c37_bench_cfg4_n2.json 2> gpurun_out/r02c37_bench_cfg4_n2.err; echo "cfg4 n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c37_bench_cfg4_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e'])
PY
