#!/bin/bash
# r02 call 63 (1 GPU): small-batch geometry with 4-CTA clusters + multicast of the ring stages (BSRNN_FUSED14_CLS=4) vs pairs:
# time-axis layer alone at B = 8, parity check, config 3 bench A/B.
mkdir -p gpurun_out
for C in 2 4; do
BSRNN_FUSED14_CLS=$C timeout 300 python tools/prof_lstm.py --fused --geo 14 --B 8 --T 1001 --K 34 --axis time --reps 3 2>&1 | grep -E "time:|co-resident|rel|Error" | tail -4
BSRNN_FUSED14_CLS=$C timeout 300 python tools/prof_lstm.py --fused --geo 14 --B 3 --T 60 --K 34 --axis time --check 2>&1 | grep -E "rel_l2|max|Error" | tail -2
done
for C in 2 4; do
BSRNN_FUSED14_CLS=$C timeout 600 python bench.py --config 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r02c63_bench_cfg3_cls$C.json 2> gpurun_out/r02c63_bench_cfg3_cls$C.err; echo "cfg3 cls=$C rc=$?"
done
python - <<'PY'
import json
for c in (2,4):
    try:
        e=json.loads(open(f'gpurun_out/r02c63_bench_cfg3_cls{c}.json').read().strip().splitlines()[-1])
        print('cls',c, round(e['ms_per_step'],2), round(e['value'],1), e['gpu_launches'], round(e['roofline']['frac'],4), {k:round(v['ms'],1) for k,v in e['per_rate'].items()})
    except Exception as ex: print('cls',c,'ERR',ex); print(open(f'gpurun_out/r02c63_bench_cfg3_cls{c}.err').read()[-800:])
PY
