#!/bin/bash
# r02 call 59 (1 GPU): where the training step's device time goes now (torch profiler, eager launches)
mkdir -p gpurun_out
timeout 600 python tools/prof_train.py > gpurun_out/r02c59_prof_train.log 2>&1; echo "prof rc=$?"
grep -A 40 "Self CUDA" gpurun_out/r02c59_prof_train.log | head -60 | cut -c 1-200
