#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -x -q -s > gpurun_out/call17_training_tests.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/call17_training_tests.log
