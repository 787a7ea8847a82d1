#!/bin/bash
# r02 call 39: racecheck after moving the ticket to its own shared-memory slot; region breakdown at small batches (8, 16, 32).
mkdir -p gpurun_out
LOG=gpurun_out/r02c39_race_small.log
: > $LOG
echo "=== racecheck sanitize_small" >> $LOG
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py 2>&1 | grep -E "rel_l2|RACECHECK SUMMARY|hazard|Error" | head -6 >> $LOG
echo "=== racecheck fused768" >> $LOG
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_fused768.py 2>&1 | grep -E "rel_l2|RACECHECK SUMMARY|hazard|Error" | head -6 >> $LOG
for b in 8 16 32; do
  timeout 600 python bench.py --batch $b --steps 4 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c39_bench_b$b.json 2> gpurun_out/r02c39_bench_b$b.err
  python - <<PY >> $LOG
import json
d=json.loads(open('gpurun_out/r02c39_bench_b$b.json').read().strip().splitlines()[-1])
print('B=$b', round(d['ms_per_step'],1), round(d['value']), round(d['e2e']['value']), d['clocks']['sm_mhz'], {k:round(v,1) for k,v in d['roofline']['regions_ms_per_step'].items()})
PY
done
cat $LOG
