#!/bin/bash
# r02 call 1 (1 GPU): full gpu test suite after the eps / invalidation changes + the recurrence slot experiment
# (how long is one step when a cluster interleaves 1, 2 or 3 sequence tiles?  B chosen so groups <= 15 clusters).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02c01_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "rel_l2|passed|failed|Error|error|eps" gpurun_out/r02c01_pytest.log | tail -40
for cfg in "64 3" "40 2" "24 1" "8 1" "8 2" "8 3"; do
  set -- $cfg
  timeout 120 python tools/prof_lstm.py --B $1 --T 1001 --K 34 --axis time --slots $2 --reps 3 2>&1 | tail -1
done
timeout 120 python tools/prof_lstm.py --B 64 --T 1001 --K 34 --axis freq --slots 3 --reps 3 2>&1 | tail -1
timeout 120 python tools/prof_lstm.py --B 64 --T 1001 --K 34 --axis freq --slots 2 --reps 3 2>&1 | tail -1
