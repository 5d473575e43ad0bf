#!/bin/bash
# round-end evidence: full GPU suite + smoke, default bench line (cfg2, with e2e / cpu baseline / fine_tune), cfg5 line,
# reference arm, launch lists (bench + fine-tune step) and full ncu captures of the GEMM kernels of a fine-tune step
mkdir -p gpurun_out
T=${TAG:-final2}
{
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
} > gpurun_out/${T}_tests.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
timeout 900 python bench.py --config cfg5 --steps 4 --warmup 3 > gpurun_out/${T}_bench_cfg5.json 2> gpurun_out/${T}_bench_cfg5.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>&1
bash tools/gpu_train_profile.sh > gpurun_out/${T}_train_breakdown.txt 2>&1
SKIP=330 COUNT=8 bash tools/gpu_ncu_train.sh > /dev/null 2>&1
cat gpurun_out/${T}_tests.log; tail -c 600 gpurun_out/${T}_bench_cfg2.json; tail -c 900 gpurun_out/${T}_bench_cfg5.json; cat gpurun_out/${T}_train_breakdown.txt | head -12
