#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_check_tc.py > gpurun_out/check_tc.log 2>&1; echo "rc=$?" >> gpurun_out/check_tc.log
tail -n 30 gpurun_out/check_tc.log
