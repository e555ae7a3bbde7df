#!/bin/bash
# ncu --set full of the step kernel for one grid/batch: bash tools/prof_grid.sh <grid> <envs> <tag>
GRID=${1:-case118}; ENVS=${2:-296}; TAG=${3:-p118}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ppn_step_kernel -s 6 -c 1 -f -o gpurun_out/prof_$TAG python bench.py --grid $GRID --envs $ENVS --steps 6 --warmup 3 --no-cpu > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$TAG.log | cut -c1-300
