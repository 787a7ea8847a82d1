#!/bin/bash
# r02 call 66 (1 GPU): Conv1d+GLU GEMM of the mask decoder walking its M tiles last-first (L2 hits on the hidden tile just written)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q -x -k "tensorcore or weight_resident or verbatim" > gpurun_out/r02c66_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02c66_pytest.log
KREG='regex:^(gemm_tc|void gemm_tc|stft|void stft)'
B="python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline --no-fp32"
for V in 1 0; do
BSRNN_GLU_REVERSE=$V timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k "$KREG" --csv --log-file gpurun_out/r02c66_ncu_launches_rev$V.csv $B > /dev/null 2>&1; echo "launch list rev=$V rc=$?"
done
