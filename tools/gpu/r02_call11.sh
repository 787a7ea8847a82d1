#!/bin/bash
# r02 call 11 (1 GPU): input-projection GEMM with A-tile multicast clusters (BSRNN_GEMM_MC=1|2|4): timing, parity, bench.
mkdir -p gpurun_out
for mc in 1 2 4; do
  BSRNN_GEMM_MC=$mc timeout 200 python tools/prof_gemm.py --which inproj --axis time --reps 4 2>&1 | tail -1
  BSRNN_GEMM_MC=$mc timeout 200 python tools/prof_gemm.py --which inproj --axis freq --reps 4 2>&1 | tail -1
done
for mc in 2 4; do
  BSRNN_GEMM_MC=$mc timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "tensorcore or fullsize or full_size or graph" 2>&1 | tail -3
done
for mc in 1 2 4; do
  BSRNN_GEMM_MC=$mc timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c11_bench_mc$mc.json 2> gpurun_out/r02c11_bench_mc$mc.err; echo "bench mc=$mc rc=$?"
  python -c "
import json
d=json.loads(open('gpurun_out/r02c11_bench_mc$mc.json').read().strip().splitlines()[-1])
r=d['roofline']['regions_ms_per_step']
print('mc=$mc', round(d['ms_per_step'],1), round(d['value']), 'inproj', round(r['inproj'],1), 'lstm', round(r['lstm_time']+r['lstm_freq'],1), 'fc', round(r['fc'],1), d['clocks']['sm_mhz'])"
done
