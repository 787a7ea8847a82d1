#!/bin/bash
# GPU call 15: CUDA-graph tests after the arena fix; ncu of the Linear+residual (epilogue 1) and input-projection
# (epilogue 4) GEMMs of a config-2 forward; graph vs host-launch bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "cuda_graph" > gpurun_out/call15_graph_tests.log 2>&1; tail -3 gpurun_out/call15_graph_tests.log
cat > /tmp/one_fwd.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
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
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:gemm_tc_kernelILi[14]E" -s 4 -c 2 -f -o gpurun_out/call15_gemm python /tmp/one_fwd.py > gpurun_out/call15_ncu.log 2>&1; tail -2 gpurun_out/call15_ncu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/call15_bench.json 2> gpurun_out/call15_bench.err; echo "bench rc=$?"; cat gpurun_out/call15_bench.json; tail -3 gpurun_out/call15_bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --graph --no-cpu-baseline > gpurun_out/call15_bench_graph.json 2>> gpurun_out/call15_bench.err; cat gpurun_out/call15_bench_graph.json
