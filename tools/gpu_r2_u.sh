#!/bin/bash
# round 2, call U: L2 row-tile prefetch in the training GEMM - A/B on the cfg5 fine-tune step
mkdir -p gpurun_out
T=${TAG:-r2u}
for pf in 0 3 1 2 0 3; do
  ZEST_GEMM_PREFETCH=$pf timeout 600 python bench.py --config cfg5 --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['fine_tune']['engines']
print('prefetch=$pf', {k[:22]: v['ms_per_step'] for k,v in e.items()}, 'opt', d['fine_tune']['with_optimizer_step']['ms_per_step'])"
done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "gradients or tc_gemm" 2>&1 | tail -2
