#!/bin/bash
# GPU call 43 (1 GPU): cp.async norm_cast kernel: parity, bench A/B against the register-staged kernel (BSRNN_PACK_SYNC=1).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/call43_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/call43_pytest_gpu.log
for v in 0 1; do
  BSRNN_PACK_SYNC=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/call43_bench_sync$v.json 2> gpurun_out/call43_bench_sync$v.err; echo "bench sync=$v rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/call43_bench_sync$v.json"))
print("sync=$v", round(d["ms_per_step"],1), "e2e", round(d["per_rank"]["e2e_ms_per_step"][0],1), d["clocks"], {k:round(x,1) for k,x in d["roofline"]["regions_ms_per_step"].items()})
PY
done
