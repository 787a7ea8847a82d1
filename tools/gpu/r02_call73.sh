#!/bin/bash
# r02 call 73 (1 GPU): last check of the final binary: smoke(), full gpu suite, default bench line
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -3
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02c73_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02c73_pytest.log
timeout 900 python bench.py > gpurun_out/r02c73_bench_cfg2.json 2> gpurun_out/r02c73_bench_cfg2.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02c73_bench_cfg2.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],2), round(d['value']), round(d['e2e']['value']), d['clocks'], round(d['roofline']['frac'],3), {k:round(v,2) for k,v in d['roofline']['regions_ms_per_step'].items()})"
