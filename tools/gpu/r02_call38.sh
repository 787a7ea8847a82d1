#!/bin/bash
# r02 call 38: compute-sanitizer on the fused schedule (small tensor-core forward, both group geometries; H = 768 kernel),
# then smoke().
mkdir -p gpurun_out
LOG=gpurun_out/r02c38_sanitizer.log
: > $LOG
for tool in memcheck racecheck synccheck; do
  echo "=== $tool sanitize_small (geo auto)" >> $LOG
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_small.py 2>&1 | grep -E "rel_l2|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error" | head -12 >> $LOG
done
echo "=== memcheck sanitize_small (geo 7)" >> $LOG
BSRNN_LSTM_FUSED_GEO=7 timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | grep -E "rel_l2|ERROR SUMMARY|Error|error" | head -8 >> $LOG
echo "=== racecheck sanitize_small (geo 7)" >> $LOG
BSRNN_LSTM_FUSED_GEO=7 timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py 2>&1 | grep -E "rel_l2|RACECHECK SUMMARY|hazard|Error|error" | head -8 >> $LOG
for tool in memcheck racecheck; do
  echo "=== $tool fused768" >> $LOG
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_fused768.py 2>&1 | grep -E "rel_l2|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error" | head -8 >> $LOG
done
cat $LOG
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -5
