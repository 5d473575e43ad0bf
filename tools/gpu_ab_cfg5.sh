#!/bin/bash
# same-box A/B of library variants on the fine-tune step: usage gpu_ab_cfg5.sh name1 name2 ... (build/variants/lib<name>.so), two rounds
mkdir -p gpurun_out; : > gpurun_out/ab_cfg5.log
for round in $(seq 1 ${ROUNDS:-2}); do for name in "$@"; do
  ZEST_B200_LIB=$PWD/build/variants/lib$name.so timeout 300 python bench.py --config cfg5 --steps ${STEPS:-5} --warmup 3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$name', round(d['ms_per_step'], 2), {k[:14]: round(v['ms_per_step'], 1) for k, v in d['fine_tune']['engines'].items()})
" >> gpurun_out/ab_cfg5.log
done; done
cat gpurun_out/ab_cfg5.log
