#!/bin/bash
# GPU call 66 (1 GPU): both directions of a BLSTM step in one launch: parity, per-step timing, FlowSE config 4.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "flowse or lstm_step" > gpurun_out/call66_pytest_flow.log 2>&1; echo "pytest rc=$?"; grep -E "FlowSE tensor|passed|failed|Error|error" gpurun_out/call66_pytest_flow.log | tail -6
P="timeout 200 python tools/prof_lstm_steps.py"
$P --N 384 --R 96 --steps 24 --check --reps 2 2>&1 | tail -3
$P --N 384 --R 1536 --steps 60 --reps 2 2>&1 | tail -1
$P --N 384 --R 40032 --steps 12 --reps 2 2>&1 | tail -1
timeout 600 python tools/bench_flowse.py --batch 32 --nfe 15 --reps 1 --precision fp16 --graph > gpurun_out/call66_flowse_fp16_graph_b32.json 2> gpurun_out/call66_flowse_b32.err; echo "rc=$?"; cat gpurun_out/call66_flowse_fp16_graph_b32.json; tail -3 gpurun_out/call66_flowse_b32.err
timeout 400 python tools/bench_flowse.py --batch 2 --nfe 15 --reps 1 --precision fp16 --graph > gpurun_out/call66_flowse_fp16_graph_b2.json 2> gpurun_out/call66_flowse_b2.err; echo "rc=$?"; cat gpurun_out/call66_flowse_fp16_graph_b2.json
