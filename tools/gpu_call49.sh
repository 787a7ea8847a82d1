#!/bin/bash
# GPU call 49 (2 GPUs): bench under torchrun exactly as the driver launches it: stdout must be ONE JSON line.
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/call49_bench_n2.json 2> gpurun_out/call49_bench_n2.err
echo "bench n2 rc=$? stdout lines: $(wc -l < gpurun_out/call49_bench_n2.json)"; cut -c1-1200 gpurun_out/call49_bench_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 \
  bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/call49_bench_ref_n2.json 2> gpurun_out/call49_bench_ref.err
echo "ref rc=$? stdout lines: $(wc -l < gpurun_out/call49_bench_ref_n2.json)"
grep -c "NCCL version" gpurun_out/call49_bench_n2.err
