#!/bin/bash
# GPU call 47 (1 GPU): final state of the session: pytest -m gpu, smoke, both bench arms (default flags), launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/call47_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/call47_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/call47_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/call47_smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/call47_bench_ref.json 2> gpurun_out/call47_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/call47_bench_ref.json
timeout 600 python bench.py > gpurun_out/call47_bench_default.json 2> gpurun_out/call47_bench_default.err; echo "bench default rc=$?"; cat gpurun_out/call47_bench_default.json
timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/call47_bench_k10.json 2> gpurun_out/call47_bench_k10.err; echo "bench k10 rc=$?"; cut -c1-800 gpurun_out/call47_bench_k10.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k 'regex:^(band_stats|gemm_f32|gemm_tc|gn_finalize|gn_stats|istft|lstm_tc|norm_cast_kb8|stft|twiddle|lstm_step|complex_mask|axpy|glu|conv5x5)' \
  --launch-skip 207 -c 207 --csv --log-file gpurun_out/call47_ncu_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/call47_ncu_bench.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/call47_ncu_launches_bench.csv
