#!/bin/bash
# GPU call 44 (2 GPUs): first multi-rank run of bench.py under torchrun (utterance sharding, no data-path collective),
# the NCCL training-step check, and the reference arm.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/call44_gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/call44_bench_n2.json 2> gpurun_out/call44_bench_n2.err
echo "bench n2 rc=$?"; cat gpurun_out/call44_bench_n2.json; tail -5 gpurun_out/call44_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  tools/ddp_check.py > gpurun_out/call44_ddp_check.log 2>&1
echo "ddp rc=$?"; grep -E "rank|rror" gpurun_out/call44_ddp_check.log | tail -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/call44_bench_ref_n2.json 2> gpurun_out/call44_bench_ref.err
echo "ref rc=$?"; cat gpurun_out/call44_bench_ref_n2.json
