#!/bin/bash
# A/B an environment toggle on the cfg2 bench: usage gpu_ab.sh VAR ; runs VAR=0,1,0,1
mkdir -p gpurun_out; : > gpurun_out/ab.log
for v in 0 1 0 1; do
  echo "== $1=$v ==" >> gpurun_out/ab.log
  env $1=$v timeout 300 python bench.py --steps 10 --warmup 3 --config cfg2 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print(round(d['value']), 'rays/s  mlp_ms', round(r['mlp_ms_per_step'], 3), 'frac', round(r['frac'], 4), r['stage_ms'])
    else:
        print(l.rstrip())
" >> gpurun_out/ab.log
done
cat gpurun_out/ab.log
