#!/bin/bash
# GPU call 41 (1 GPU): end-of-session record: pytest -m gpu, smoke, bench (both arms), ncu launch list of one forward of
# the bench workload (our kernels only, second forward), ncu --set full of the three dominant kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/call41_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/call41_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/call41_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/call41_smoke.log
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/call41_bench_ref.json 2> gpurun_out/call41_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/call41_bench_ref.json
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/call41_bench.json 2> gpurun_out/call41_bench.err; echo "bench rc=$?"; cat gpurun_out/call41_bench.json; tail -3 gpurun_out/call41_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bsrnn --launch-skip 1035 -c 1035 --csv --log-file gpurun_out/call41_ncu_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/call41_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_kernel -c 1 -o gpurun_out/call41_lstm_v8_freq \
  python tools/prof_lstm.py --B 64 --T 1001 --K 34 --axis freq --slots 3 --reps 1 > gpurun_out/call41_ncu_lstm.log 2>&1; echo "ncu lstm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 1 -o gpurun_out/call41_gemm_inproj \
  python tools/prof_gemm.py --which inproj --axis time --reps 1 --nobias > gpurun_out/call41_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 1 -o gpurun_out/call41_gemm_fc \
  python tools/prof_gemm.py --which fc --axis time --reps 1 > gpurun_out/call41_ncu_gemm_fc.log 2>&1; echo "ncu gemm fc rc=$?"
ls -la gpurun_out | tail -14
