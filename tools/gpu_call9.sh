#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./build/ubench > gpurun_out/call9_ubench.log 2>&1
cat gpurun_out/call9_ubench.log
