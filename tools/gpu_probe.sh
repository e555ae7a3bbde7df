#!/bin/bash
# Quick probe on the GPU box: bench (short), phase timing and one full ncu capture with source for the line breakdown.
# Usage (under gpurun): bash tools/gpu_probe.sh <tag> [grid] [envs]
TAG=${1:-p}; GRID=${2:-case14}; ENVS=${3:-4096}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python bench.py --grid $GRID --envs $ENVS --steps 100 --warmup 10 --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cut -c1-1500 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
timeout 300 python tools/phase_timing.py $GRID $ENVS > $OUT/phase_$TAG.log 2>&1; echo "phase rc=$?"; tail -12 $OUT/phase_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ppn_step_kernel -s 10 -c 1 -f -o $OUT/prof_$TAG python bench.py --grid $GRID --envs $ENVS --steps 12 --warmup 3 --no-cpu > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT | head -30
