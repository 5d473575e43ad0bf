#!/bin/bash
# round 2, call D (2 GPUs): dynamic tile scheduler - GPU suite, 1-GPU A/B (static vs dynamic), 2-GPU strong scaling (both transports)
mkdir -p gpurun_out
T=${TAG:-r2d}
N=${NGPU:-2}
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/${T}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_tests.log
tail -3 gpurun_out/${T}_tests.log
for mode in 0 1 0 1; do
  ZEST_TC_STATIC=$mode timeout 300 python bench.py --steps 10 --warmup 3 --no-fine-tune --no-cpu-baseline --no-e2e > gpurun_out/${T}_ab_static${mode}.json 2>> gpurun_out/${T}_ab.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/${T}_ab_static${mode}.json').read().strip().splitlines()[-1])
print('static=$mode', round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks'].get('power_w'))
"
done
for tr in ipc nccl; do
  ZEST_FRAME_TRANSPORT=$tr timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench_${N}gpu_${tr}.json 2> gpurun_out/${T}_bench_${N}gpu_${tr}.err
  echo "bench $tr rc=$?"
  python - "gpurun_out/${T}_bench_${N}gpu_${tr}.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","scaling","sharded_frame_equals_single_gpu","steps_ms")})
    print("e2e", d["e2e"]["value"] if d.get("e2e") else None, "pose_parallel", d["pose_parallel_weak"]["value"], "cfg3", d["cfg3_strong"])
except Exception as e:
    print("no line:", e)
PY
done
