#!/bin/bash
# GPU call 14: ncu of the input-projection (epilogue 4) and Linear+residual (epilogue 1) GEMMs inside a config-2 forward;
# compute-sanitizer on the failing fp32 CUDA-graph test.
mkdir -p gpurun_out
cat > /tmp/one_fwd.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from oracle import restated as R
from urgent2026_challenge_track1_b200 import BSRNN_SE
torch.manual_seed(0)
m = BSRNN_SE(196, 1, precision="fp16").cuda()
B, n = 64, 480000
x = (0.1 * torch.randn(B, n)).cuda()
lens = torch.full((B,), n, dtype=torch.int32)
for _ in range(2):
    m(x, lens, 48000)
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 6 -c 2 -f -o gpurun_out/call14_gemm python /tmp/one_fwd.py > gpurun_out/call14_ncu.log 2>&1; tail -2 gpurun_out/call14_ncu.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -k "cuda_graph and fp32" > gpurun_out/call14_sanitizer.log 2>&1; grep -E "Invalid|at |by thread|Address|ERROR SUMMARY|passed|failed" gpurun_out/call14_sanitizer.log | head -30
python - <<'PY'
import torch
x = torch.empty(14_500_000_000 // 2, dtype=torch.float16, device="cuda")
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); x.zero_(); e1.record(); torch.cuda.synchronize()
    print("memset 14.5 GB:", e0.elapsed_time(e1), "ms ->", 14.5 / e0.elapsed_time(e1), "TB/s")
PY
