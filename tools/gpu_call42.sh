#!/bin/bash
# GPU call 42 (1 GPU): bench with the reordered setup (capture first, 3 eager region passes, warm-up, timed loop), and the
# ncu launch list of ONE forward of the bench workload (our kernels; the first eager forward is skipped).
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/call42_bench.json 2> gpurun_out/call42_bench.err; echo "bench rc=$?"; cat gpurun_out/call42_bench.json; tail -3 gpurun_out/call42_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/call42_bench_k10.json 2> gpurun_out/call42_bench_k10.err; echo "bench k10 rc=$?"; cut -c1-700 gpurun_out/call42_bench_k10.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k 'regex:^(band_stats|gemm_f32|gemm_tc|gn_finalize|gn_stats|istft|lstm_tc|norm_cast_kb8|stft|twiddle|lstm_step|complex_mask|axpy|glu|conv5x5)' \
  --launch-skip 207 -c 207 --csv --log-file gpurun_out/call42_ncu_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/call42_ncu_bench.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/call42_ncu_launches_bench.csv
