#!/bin/bash
# r02 call 14: flag-group recurrence with the cell state of slots 2/3 in TMEM: numerics with 3 and 4 slots, timing.
for sl in 3 4; do
  timeout 120 python tools/prof_lstm.py --B 24 --T 41 --K 34 --axis freq --flag --slots $sl --check --reps 1 2>&1 | grep -E "CHECK|rror" | tail -1
  timeout 120 python tools/prof_lstm.py --B 24 --T 61 --K 34 --axis time --flag --slots $sl --check --reps 1 2>&1 | grep -E "CHECK|rror" | tail -1
done
for args in "--axis time --slots 2" "--axis time --slots 3" "--axis freq --slots 3" "--axis freq --slots 4" "--axis time --slots 4"; do
  timeout 120 python tools/prof_lstm.py --B 64 --T 1001 --K 34 $args --flag --reps 3 2>&1 | tail -1
done
