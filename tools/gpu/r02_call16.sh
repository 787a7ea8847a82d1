#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02c16_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02c16_pytest.log
