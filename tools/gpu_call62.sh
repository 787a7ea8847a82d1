#!/bin/bash
# GPU call 62 (1 GPU): step-wise tensor-core BLSTM in the FlowSE path: parity tests, FlowSE timing f32 vs fp16 dual path.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "flowse or lstm_step" > gpurun_out/call62_pytest_flow.log 2>&1; echo "pytest rc=$?"; grep -E "FlowSE tensor|passed|failed|Error|error" gpurun_out/call62_pytest_flow.log | tail -8
timeout 300 python tools/bench_flowse.py --batch 2 --nfe 15 --reps 1 --precision fp16 > gpurun_out/call62_flowse_fp16_b2.json 2> gpurun_out/call62_flowse_fp16_b2.err; echo "rc=$?"; cat gpurun_out/call62_flowse_fp16_b2.json; tail -2 gpurun_out/call62_flowse_fp16_b2.err
timeout 400 python tools/bench_flowse.py --batch 8 --nfe 15 --reps 1 --precision fp16 > gpurun_out/call62_flowse_fp16_b8.json 2> gpurun_out/call62_flowse_fp16_b8.err; echo "rc=$?"; cat gpurun_out/call62_flowse_fp16_b8.json; tail -2 gpurun_out/call62_flowse_fp16_b8.err
