#!/bin/bash
# round 2, call B (2 GPUs): full GPU suite, then the ray-sharded strong-scaling bench with both frame transports
mkdir -p gpurun_out
T=${TAG:-r2b}
N=${NGPU:-2}
timeout 1500 python -m pytest tests -q -m gpu --durations=8 -s > gpurun_out/${T}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_tests.log
for tr in ipc nccl; do
  ZEST_FRAME_TRANSPORT=$tr timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench_${N}gpu_${tr}.json 2> gpurun_out/${T}_bench_${N}gpu_${tr}.err
  echo "bench $tr rc=$?"
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-fine-tune > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
grep -E "passed|failed|rc=" gpurun_out/${T}_tests.log | tail -5
grep -E "^FAILED|Error" gpurun_out/${T}_tests.log | head -20
for f in gpurun_out/${T}_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","scaling","sharded_frame_equals_single_gpu","pose_parallel_weak","cfg3_strong","parity")})
    print("e2e", d["e2e"]["value"] if d.get("e2e") else None, d["config"]["parallelism"], "frac", d["roofline"]["frac"], d["roofline"]["whole_step_frac"])
except Exception as e:
    print("no line:", e)
PY
tail -5 ${f%.json}.err
done
