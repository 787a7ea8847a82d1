#!/bin/bash
mkdir -p gpurun_out
python tools/prof_lstm.py 16 401 34 time 3 > gpurun_out/prof_lstm.log 2>&1
python tools/prof_lstm.py 3 401 34 time 3 >> gpurun_out/prof_lstm.log 2>&1
python tools/prof_lstm.py 16 401 34 time 3 1 >> gpurun_out/prof_lstm.log 2>&1
python tools/prof_lstm.py 4 401 34 freq 3 >> gpurun_out/prof_lstm.log 2>&1
cat gpurun_out/prof_lstm.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lstm_tc -c 1 -o gpurun_out/prof_lstm -f python tools/prof_lstm.py 3 401 34 time 1 > gpurun_out/ncu_lstm.log 2>&1
tail -n 3 gpurun_out/ncu_lstm.log
ls -la gpurun_out/*.ncu-rep
