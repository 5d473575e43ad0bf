#!/bin/bash
# round 2, call N (2 GPUs): the ipc transport with a direct peer cudaMemcpyAsync - timeline probe + bench A/B against nccl
mkdir -p gpurun_out
T=${TAG:-r2n}
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/scale_probe.py > gpurun_out/${T}_probe.log 2>&1
grep -E "stand-alone|ipc\] rank" -A1 gpurun_out/${T}_probe.log | cut -c1-900 | tail -8
for tr in ipc nccl ipc nccl; do
  ZEST_FRAME_TRANSPORT=$tr timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/${T}_bench_${N}gpu_${tr}.json 2> gpurun_out/${T}_bench_${N}gpu_${tr}.err
  python - "gpurun_out/${T}_bench_${N}gpu_${tr}.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["config"]["parallelism"][-40:], round(d["value"]), d["ms_per_step"], d["sharded_frame_equals_single_gpu"], d["steps_ms"][:6])
except Exception as e:
    print("no line:", e)
PY
done
