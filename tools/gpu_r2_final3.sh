#!/bin/bash
# round 2, end of session: full GPU suite, smoke, default bench line (cfg2), cfg5 line
mkdir -p gpurun_out
T=r2final3
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
timeout 600 python bench.py --config cfg5 --steps 5 --warmup 3 > gpurun_out/${T}_bench_cfg5.json 2> gpurun_out/${T}_bench_cfg5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2final3_bench_cfg2.json').read().strip().splitlines()[-1])
print(round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', d['roofline']['frac'], d['roofline'].get('whole_step_frac'), 'clocks', d.get('clocks'))
print('fine_tune', {k: v['ms_per_step'] for k, v in d.get('fine_tune', {}).get('engines', {}).items()})
print('pipeline', d['next_rows']['full_frame_pipeline'])
d=json.loads(open('gpurun_out/r2final3_bench_cfg5.json').read().strip().splitlines()[-1])
print('cfg5', d['value'], d['ms_per_step'], {k: v['ms_per_step'] for k, v in d['fine_tune']['engines'].items()}, d['fine_tune']['with_optimizer_step']['ms_per_step'])
PY
tail -2 gpurun_out/${T}_bench_cfg2.err
