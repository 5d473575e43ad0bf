#!/bin/bash
# repeatability: N back-to-back default bench runs, print value + stage times
for i in $(seq 1 ${1:-6}); do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], d['clocks']['samples'], d['roofline']['stage_ms'], d['steps_ms'])"
done
