#!/bin/bash
# GPU call 70 (1 GPU): FlowSE condition_fc on tensor cores: flow tests, fp16-vs-f32 at full width, config 4 timing.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "flowse or lstm_step" > gpurun_out/call70_pytest_flow.log 2>&1; echo "pytest rc=$?"; grep -E "FlowSE tensor|passed|failed|Error|error" gpurun_out/call70_pytest_flow.log | tail -6
timeout 600 python tools/flowse_fp16_vs_f32.py 2>&1 | tail -1
timeout 600 python tools/bench_flowse.py --batch 32 --nfe 15 --reps 1 --graph > gpurun_out/call70_flowse_b32.json 2> gpurun_out/call70_flowse_b32.err; echo "rc=$?"; cat gpurun_out/call70_flowse_b32.json; tail -3 gpurun_out/call70_flowse_b32.err
