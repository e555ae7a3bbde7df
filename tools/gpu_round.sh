#!/bin/bash
# One GPU-box visit: parity tests, bench (N=1), ncu launch list and one full capture of the step kernel.
# Usage (under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/gputests_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/gputests_$TAG.log
timeout 600 python bench.py --steps 200 --warmup 20 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err; echo "ref rc=$?"; cat $OUT/bench_ref_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/bench_under_ncu_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ppn_step_kernel -s 10 -c 1 -f -o $OUT/prof_$TAG python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
