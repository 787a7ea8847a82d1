#!/bin/bash
# r02 call 12: FlowSE step kernel (H=768): per-step time under CUDA-graph replay vs N-tile width, time-axis (R=1536) and
# band-axis (R=40032) shapes of config 4.
for bn in 256 192 128 96; do
  timeout 120 python tools/prof_lstm_steps.py --R 1536 --steps 200 --N 384 --graph --bn $bn --reps 2 2>&1 | tail -1
done
for bn in 256 128; do
  timeout 200 python tools/prof_lstm_steps.py --R 40032 --steps 24 --N 384 --graph --bn $bn --reps 2 2>&1 | tail -1
done
timeout 120 python tools/prof_lstm_steps.py --R 1536 --steps 200 --N 384 --reps 2 2>&1 | tail -1
