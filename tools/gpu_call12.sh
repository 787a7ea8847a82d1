#!/bin/bash
# GPU call 12: v4 recurrence after the per-slot acc_full fix: parity, probes, ncu stall picture, bench.
mkdir -p gpurun_out
LOG=gpurun_out/call12_lstm_v4.log
: > $LOG
for s in 1 2 3; do
  timeout 120 python tools/prof_lstm.py --B 12 --T 40 --K 34 --axis time --slots $s --check --reps 1 >> $LOG 2>&1 || echo "FAILED time slots=$s rc=$?" >> $LOG
  timeout 120 python tools/prof_lstm.py --B 3 --T 300 --K 34 --axis freq --slots $s --check --reps 1 >> $LOG 2>&1 || echo "FAILED freq slots=$s rc=$?" >> $LOG
done
timeout 120 python tools/prof_lstm.py --B 40 --T 60 --K 34 --axis time --slots 2 --maxcl 3 --check --reps 1 >> $LOG 2>&1 || echo "FAILED multi-group time" >> $LOG
timeout 120 python tools/prof_lstm.py --B 40 --T 60 --K 34 --axis freq --slots 3 --maxcl 2 --check --reps 1 >> $LOG 2>&1 || echo "FAILED multi-group freq" >> $LOG
P="timeout 120 python tools/prof_lstm.py --B 64 --T 201 --K 34 --reps 2"
for s in 1 2 3; do
  $P --axis time --slots $s --trace --trace-cid 1 >> $LOG 2>&1 || echo "FAILED T201 time $s" >> $LOG
  $P --axis freq --slots $s --trace --trace-cid 1 >> $LOG 2>&1 || echo "FAILED T201 freq $s" >> $LOG
done
timeout 200 python tools/prof_lstm.py --B 64 --T 1001 --K 34 --reps 2 --axis time --slots 3 >> $LOG 2>&1
timeout 200 python tools/prof_lstm.py --B 64 --T 1001 --K 34 --reps 2 --axis freq --slots 3 >> $LOG 2>&1
grep -E "CHECK|FAILED|ms,|producer|mma  |epilogue|rror" $LOG | tail -70
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_kernel -c 1 -f -o gpurun_out/call12_lstm_v4 \
  python tools/prof_lstm.py --B 64 --T 41 --K 34 --axis time --slots 3 --reps 1 > gpurun_out/call12_ncu.log 2>&1; tail -1 gpurun_out/call12_ncu.log
if [ $(grep -c "CHECK.*OK" $LOG) -ge 8 ] && ! grep -q "FAIL" $LOG; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/call12_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/call12_pytest_gpu.log
  timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/call12_bench.json 2> gpurun_out/call12_bench.err; echo "bench rc=$?"; cat gpurun_out/call12_bench.json
fi
