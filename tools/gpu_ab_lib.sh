#!/bin/bash
# same-box A/B of library variants and env toggles, two rounds.
# usage: gpu_ab_lib.sh "name[,VAR=val,...]" ...      (build/variants/lib<name>.so)
mkdir -p gpurun_out; : > gpurun_out/ab.log
for round in 1 2; do for spec in "$@"; do
  name=${spec%%,*}; envs=$(echo "${spec#$name}" | tr ',' ' ')
  echo "== $spec ==" >> gpurun_out/ab.log
  env $envs ZEST_B200_LIB=$PWD/build/variants/lib$name.so timeout 300 python bench.py --steps 10 --warmup 3 --config cfg2 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print(round(d['value']), 'rays/s  step_ms', round(d['ms_per_step'], 3), 'mlp_ms', round(r['mlp_ms_per_step'], 3), 'frac', round(r['frac'], 4), r['stage_ms'])
    else:
        print(l.rstrip())
" >> gpurun_out/ab.log
done; done
cat gpurun_out/ab.log
