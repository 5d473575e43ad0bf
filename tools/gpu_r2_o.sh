#!/bin/bash
# round 2, call O (8 GPUs): frame transports A/B at N = 8 (ipc with direct peer memcpy vs nccl)
mkdir -p gpurun_out
T=${TAG:-r2o}
N=8
for tr in ipc nccl ipc nccl; do
  ZEST_FRAME_TRANSPORT=$tr timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus $N --steps 20 --warmup 3 --no-e2e > gpurun_out/${T}_bench_${N}gpu_${tr}.json 2> gpurun_out/${T}_bench_${N}gpu_${tr}.err
  python - "gpurun_out/${T}_bench_${N}gpu_${tr}.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["config"]["parallelism"][-40:], round(d["value"]), d["ms_per_step"], d["sharded_frame_equals_single_gpu"], d["steps_ms"][:4], d["steps_ms"][-3:], "cfg3", round(d["cfg3_strong"]["value"]))
except Exception as e:
    print("no line:", e)
PY
done
