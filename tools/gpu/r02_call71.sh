#!/bin/bash
# r02 call 71 (1 GPU): memcheck of the final tree: the small forward, the mid-size forward that reaches the weight-resident /
# specialised GEMM schedules and the tensor-map paths at several tiles per CTA, one fused training block
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/r02c71_memcheck_small.log 2>&1; echo "small rc=$?"; tail -2 gpurun_out/r02c71_memcheck_small.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "weight_resident_gemm_paths" > gpurun_out/r02c71_memcheck_midsize.log 2>&1; echo "midsize rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02c71_memcheck_midsize.log | tail -3
