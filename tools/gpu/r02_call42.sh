#!/bin/bash
# r02 call 42: the advertised A/B switches still work: parity suites under the unfused schedule, forced geometries, and the
# step-wise FlowSE path.
mkdir -p gpurun_out
LOG=gpurun_out/r02c42_switches.log
: > $LOG
t() { echo "=== $1" >> $LOG; shift; env "$@" timeout 900 python -m pytest tests -m gpu -q -x -k "parity or reference or fullsize" 2>&1 | tail -2 >> $LOG; }
t "BSRNN_LSTM_FUSED=none" BSRNN_LSTM_FUSED=none
t "BSRNN_LSTM_FUSED=freq" BSRNN_LSTM_FUSED=freq
t "BSRNN_LSTM_FUSED_GEO=8" BSRNN_LSTM_FUSED_GEO=8
t "BSRNN_LSTM_FUSED_GEO=7" BSRNN_LSTM_FUSED_GEO=7
t "BSRNN_LSTM_FUSED_GEO=14" BSRNN_LSTM_FUSED_GEO=14
t "BSRNN_FLOWSE_FUSED=0" BSRNN_FLOWSE_FUSED=0
t "BSRNN_LSTM_SCHED=cluster BSRNN_LSTM_FUSED=none" BSRNN_LSTM_SCHED=cluster BSRNN_LSTM_FUSED=none
cat $LOG
