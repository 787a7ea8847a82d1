#!/bin/bash
# r02 call 13: FlowSE step kernel with the W_hh tile resident + h tiles multicast (BN = 96): parity, per-step time.
for R in 150 1536; do timeout 200 python tools/prof_lstm_steps.py --R $R --steps 9 --N 384 --check --reps 1 2>&1 | tail -1; done
for mc in 1 2 4; do
  BSRNN_STEP_MC=$mc timeout 120 python tools/prof_lstm_steps.py --R 1536 --steps 200 --N 384 --graph --reps 2 2>&1 | tail -1
  BSRNN_STEP_MC=$mc timeout 200 python tools/prof_lstm_steps.py --R 40032 --steps 24 --N 384 --graph --reps 2 2>&1 | tail -1
done
BSRNN_STEP_RESIDENT=0 timeout 120 python tools/prof_lstm_steps.py --R 1536 --steps 200 --N 384 --graph --reps 2 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q -x -k "flowse or lstm_step" 2>&1 | tail -3
