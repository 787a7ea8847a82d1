#!/bin/bash
# GPU call 64 (1 GPU): FlowSE network evaluation replayed from a CUDA graph (step-wise tensor-core mode).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "flowse" > gpurun_out/call64_pytest_flow.log 2>&1; echo "pytest rc=$?"; grep -E "FlowSE tensor|passed|failed|Error|error" gpurun_out/call64_pytest_flow.log | tail -8
timeout 400 python tools/bench_flowse.py --batch 2 --nfe 15 --reps 1 --precision fp16 --graph > gpurun_out/call64_flowse_fp16_graph_b2.json 2> gpurun_out/call64_flowse_b2.err; echo "rc=$?"; cat gpurun_out/call64_flowse_fp16_graph_b2.json; tail -3 gpurun_out/call64_flowse_b2.err
timeout 600 python tools/bench_flowse.py --batch 32 --nfe 15 --reps 1 --precision fp16 --graph > gpurun_out/call64_flowse_fp16_graph_b32.json 2> gpurun_out/call64_flowse_b32.err; echo "rc=$?"; cat gpurun_out/call64_flowse_fp16_graph_b32.json; tail -3 gpurun_out/call64_flowse_b32.err
