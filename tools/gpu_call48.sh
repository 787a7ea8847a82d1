#!/bin/bash
# GPU call 48: where do the occasional ~200 ms stalls in one of the two timed loops come from?  bench x4 with the NVML
# clock sampler and x4 without (BSRNN_BENCH_NO_NVML=1), K=5.
mkdir -p gpurun_out
for i in 1 2 3 4; do
  for nv in 0 1; do
    BSRNN_BENCH_NO_NVML=$nv timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/call48_b_${nv}_$i.json 2>/dev/null
    python - <<PY
import json
d=json.load(open("gpurun_out/call48_b_${nv}_$i.json"))
print("no_nvml=$nv run $i: value loop", round(d["ms_per_step"],1), "ms  e2e loop", round(d["per_rank"]["e2e_ms_per_step"][0],1), "ms", d["clocks"]["sm_mhz"], d["clocks"].get("sm_mhz_e2e"))
PY
  done
done
