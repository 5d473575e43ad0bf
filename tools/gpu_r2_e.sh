#!/bin/bash
# round 2, call E (8 GPUs): ray-sharded strong scaling at N = 8 (and 4)
mkdir -p gpurun_out
T=${TAG:-r2e}
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err
  echo "bench N=$N rc=$?"
  python - "gpurun_out/${T}_bench_${N}gpu.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","scaling","sharded_frame_equals_single_gpu","steps_ms","host_enqueue_ms_per_step")})
    print("e2e", d["e2e"]["value"] if d.get("e2e") else None, "pose_parallel", d["pose_parallel_weak"], "cfg3", d["cfg3_strong"])
    print(d["roofline"]["stage_ms"], d["clocks"])
except Exception as e:
    print("no line:", e)
PY
  tail -3 gpurun_out/${T}_bench_${N}gpu.err
done
