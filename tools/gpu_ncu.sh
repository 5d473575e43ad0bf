#!/bin/bash
# ncu evidence: per-launch device times of one bench command + one full capture of the MLP kernel.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --config cfg2 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 6 -c 2 -o gpurun_out/prof_mlp $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -8
tail -2 gpurun_out/ncu_full.log
