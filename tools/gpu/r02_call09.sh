#!/bin/bash
# r02 call 9 (1 GPU): full gpu suite + smoke, graphed training step, config 5 (graph / host launches).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02c09_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02c09_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for extra in "" "--no-graph"; do
  timeout 600 python bench.py --config 5 $extra > gpurun_out/r02c09_cfg5.json 2> gpurun_out/r02c09_cfg5.err; echo "cfg5 $extra rc=$?"; tail -2 gpurun_out/r02c09_cfg5.err
  python -c "
import json
d=json.loads(open('gpurun_out/r02c09_cfg5.json').read().strip().splitlines()[-1])
print('$extra', d['ms_per_step'], d['value'], d['e2e']['value'], d['split_ms'], d['loss'], d['impl_notes']['launch'][:30])"
  [ -z "$extra" ] && cp gpurun_out/r02c09_cfg5.json gpurun_out/r02c09_cfg5_graph.json
done
