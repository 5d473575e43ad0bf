#!/bin/bash
# round 2, call A: full GPU suite (incl. the BASELINE-shape parity and reference-caller tests), default bench line,
# and ncu --set full of the stand-alone stage kernels (gather fwd / bwd, composites, ray builder)
mkdir -p gpurun_out
T=${TAG:-r2a}
timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 -s > gpurun_out/${T}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
timeout 300 python tools/stage_step.py > gpurun_out/${T}_stage_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gather_fwd|gather_bwd|composite|build_rays' -s 6 -c 6 \
    -o gpurun_out/${T}_stages python tools/stage_step.py > gpurun_out/${T}_stage_ncu.log 2>&1
tail -25 gpurun_out/${T}_tests.log; tail -3 gpurun_out/${T}_smoke.log; tail -c 1500 gpurun_out/${T}_bench_cfg2.json; tail -3 gpurun_out/${T}_stage_ncu.log
