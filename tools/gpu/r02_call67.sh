#!/bin/bash
# r02 call 67 (1 GPU): A/B in the real (warm-L2, graph-replayed) step: Conv1d+GLU GEMM tile order, evict_last policy on the hidden tile
mkdir -p gpurun_out
run() { env "$@" timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fp32 2>/dev/null | tail -1 > gpurun_out/r02c67_tmp.json; python - "$*" <<'PY'
import json, sys
d=json.loads(open('gpurun_out/r02c67_tmp.json').read())
print(sys.argv[1], round(d['ms_per_step'],2), d['clocks']['sm_mhz'], {k:round(v,2) for k,v in d['roofline']['regions_ms_per_step'].items()})
PY
}
run BSRNN_GLU_REVERSE=1
run BSRNN_GLU_REVERSE=0
run BSRNN_GLU_REVERSE=1 BSRNN_HIDDEN_L2=1
run BSRNN_GLU_REVERSE=0
run BSRNN_GLU_REVERSE=1
