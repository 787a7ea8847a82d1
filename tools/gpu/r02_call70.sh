#!/bin/bash
# r02 call 70 (1 GPU): verification of the committed tree: smoke(), full gpu suite, default bench line (all legs), reference arm,
# configs 3 / 4 / 5, launch list of one config-2 step.
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -4
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02c70_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02c70_pytest.log
S=$(date +%s)
timeout 1200 python bench.py > gpurun_out/r02c70_bench_cfg2.json 2> gpurun_out/r02c70_bench_cfg2.err; echo "cfg2 rc=$? wall=$(( $(date +%s) - S )) s"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02c70_bench_reference.json 2>/dev/null; echo "reference rc=$?"; tail -1 gpurun_out/r02c70_bench_reference.json | cut -c 1-300
for C in 3 4 5; do
timeout 1200 python bench.py --config $C --no-cpu-baseline --no-library-baseline > gpurun_out/r02c70_bench_cfg$C.json 2> gpurun_out/r02c70_bench_cfg$C.err; echo "cfg$C rc=$?"
done
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c70_bench_cfg2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','scaling','dtype','gpu_launches')}, d['e2e'], round(d['roofline']['frac'],3), {k:(round(v['ms_per_launch'],2), round(v['frac'],3)) for k,v in d['roofline']['per_axis'].items()}, d['roofline']['whole_step'])
print('cpu', d.get('cpu_baseline')); print('lib', {k:(round(v['value']) if isinstance(v,dict) else v) for k,v in d.get('library_baseline',{}).items() if k in ('tf32','fp16_autocast','fp32')}); print('fp32', d.get('fp32_mode',{}).get('value')); print(d['clocks'])
print({k:(round(v['achieved']),round(v['frac'],2)) for k,v in d['roofline']['other_kernels'].items()})
print({k:round(v,2) for k,v in d['roofline']['regions_ms_per_step'].items()})
for c in (3,4,5):
    try:
        e=json.loads(open(f'gpurun_out/r02c70_bench_cfg{c}.json').read().strip().splitlines()[-1])
        print('cfg',c, round(e['ms_per_step'],2), round(e['value'],1), e['e2e']['value'], e['clocks']['sm_mhz'], e.get('roofline',{}).get('frac'))
    except Exception as ex: print('cfg',c,'ERR',ex)
PY
KREG='regex:^(lstm_|gemm_|norm_cast|istft|stft|band_|gn_|void lstm|void gemm|void stft|void norm)'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" --csv --log-file gpurun_out/r02c70_ncu_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-library-baseline --no-fp32 > gpurun_out/r02c70_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
