#!/bin/bash
# GPU call 10: ncu source-level capture of the v3 recurrence kernel (slots=1 and slots=3) to find the epilogue stalls.
mkdir -p gpurun_out
for s in 1 3; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_kernel -c 1 -f -o gpurun_out/call10_lstm_s$s \
  python tools/prof_lstm.py --B 64 --T 41 --K 34 --axis time --slots $s --variant 0 --reps 1 > gpurun_out/call10_ncu_s$s.log 2>&1
tail -2 gpurun_out/call10_ncu_s$s.log
done
ls -la gpurun_out/*.ncu-rep
