#!/bin/bash
# round 2, final single-GPU evidence: full GPU suite + smoke, default bench line, reference arm, cfg3 / cfg1 / cfg5 lines
mkdir -p gpurun_out
T=${TAG:-r2final}
{
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
} > gpurun_out/${T}_tests.log 2>&1
cat gpurun_out/${T}_tests.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
timeout 600 python bench.py --config cfg3 --steps 10 --warmup 3 --no-fine-tune > gpurun_out/${T}_bench_cfg3.json 2> gpurun_out/${T}_bench_cfg3.err
timeout 600 python bench.py --config cfg1 --steps 10 --warmup 3 --no-fine-tune > gpurun_out/${T}_bench_cfg1.json 2> gpurun_out/${T}_bench_cfg1.err
timeout 900 python bench.py --config cfg5 --steps 4 --warmup 3 > gpurun_out/${T}_bench_cfg5.json 2> gpurun_out/${T}_bench_cfg5.err
python - <<'PY'
import json
for c in ("ref","cfg2","cfg3","cfg1","cfg5"):
    try:
        d=json.loads(open(f'gpurun_out/r2final_bench_{c}.json').read().strip().splitlines()[-1])
        print(c, round(d['value'],1), 'e2e', d.get('e2e',{}).get('value') if d.get('e2e') else None, 'frac', d.get('roofline',{}).get('frac') if d.get('roofline') else None,
              'cpu', d.get('cpu_baseline'), 'parity', d.get('parity'))
        if c=='cfg2': print(json.dumps(d['next_rows']['f3_mvsnet']), d['roofline']['whole_step_frac'], d['clocks'])
    except Exception as e:
        print(c, 'no line', e)
PY
