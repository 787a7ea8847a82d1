#!/bin/bash
# GPU call 18 (2 GPUs): NCCL check of the training step, train-step timing at 1 and 2 GPUs, inference bench at 2 GPUs.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tools/ddp_check.py > gpurun_out/call18_ddp_check.log 2>&1; echo "ddp rc=$?"; grep -E "rank|rror" gpurun_out/call18_ddp_check.log | tail -6
timeout 600 python tools/bench_train.py > gpurun_out/call18_train_1gpu.json 2> gpurun_out/call18_train.err; cat gpurun_out/call18_train_1gpu.json; tail -3 gpurun_out/call18_train.err
timeout 600 $TR tools/bench_train.py > gpurun_out/call18_train_2gpu.json 2>> gpurun_out/call18_train.err; cat gpurun_out/call18_train_2gpu.json
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/call18_bench_2gpu.json 2> gpurun_out/call18_bench.err; echo "bench2 rc=$?"; cat gpurun_out/call18_bench_2gpu.json; tail -3 gpurun_out/call18_bench.err
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/call18_bench_ref.json 2>> gpurun_out/call18_bench.err; cat gpurun_out/call18_bench_ref.json
