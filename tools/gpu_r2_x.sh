#!/bin/bash
# round 2, call X (8 GPUs): BASELINE config 4 - 1080p wander-path frames, each frame ray-sharded across the ranks
mkdir -p gpurun_out
T=${TAG:-r2x}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus 8 --config cfg4 --steps 60 --warmup 3 > gpurun_out/${T}_cfg4_8gpu.json 2> gpurun_out/${T}_cfg4_8gpu.err
echo "rc=$?"; tail -c 1200 gpurun_out/${T}_cfg4_8gpu.json; tail -3 gpurun_out/${T}_cfg4_8gpu.err
timeout 600 python bench.py --config cfg4 --steps 10 --warmup 3 > gpurun_out/${T}_cfg4_1gpu.json 2> gpurun_out/${T}_cfg4_1gpu.err
echo "rc=$?"; tail -c 600 gpurun_out/${T}_cfg4_1gpu.json
