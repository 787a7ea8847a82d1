#!/bin/bash
# r02 call 55 (1 GPU): bisect the full-size mismatch (rows of a 64-utterance batch vs the same utterances in a batch of 2)
mkdir -p gpurun_out
run() { env "$@" timeout 300 python tools/debug_fullsize.py 10 64 2>&1 | grep -E "switches|Error|error" | tail -2; }
run A=0
run BSRNN_FC_TMAP=0 BSRNN_PACK_TMAP=0
run BSRNN_FC_TMAP=0
run BSRNN_PACK_TMAP=0
run BSRNN_BAND_SPLIT_TC=0
run BSRNN_BAND_SPLIT_GROUPED=0
run BSRNN_PACK_ROWS=128
run BSRNN_FC_RUNS=0 BSRNN_PACK_RUNS=0
run BSRNN_TANH_BULK=0
run BSRNN_MASKDEC_SHARED_NORM=0
run BSRNN_FC_EPI=ldst
timeout 300 python tools/debug_fullsize.py 10 16 2>&1 | grep switches
timeout 300 python tools/debug_fullsize.py 3 64 2>&1 | grep switches
