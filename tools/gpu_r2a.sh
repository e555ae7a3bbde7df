#!/bin/bash
# Round 2, first visit: the new parity tests on the unchanged library + baselines of every BASELINE configuration.
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -s > $OUT/gputests_r2a.log 2>&1; echo "tests rc=$?"; tail -15 $OUT/gputests_r2a.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu > $OUT/bench_r2a_14.json 2> $OUT/bench_r2a.err; echo "bench rc=$?"; cat $OUT/bench_r2a_14.json
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --grid case30 --envs 8192 > $OUT/bench_r2a_30.json 2>> $OUT/bench_r2a.err; cat $OUT/bench_r2a_30.json
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --grid case118 --envs 8192 > $OUT/bench_r2a_118.json 2>> $OUT/bench_r2a.err; cat $OUT/bench_r2a_118.json
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --grid case118 --envs 4096 --agent random > $OUT/bench_r2a_118r.json 2>> $OUT/bench_r2a.err; cat $OUT/bench_r2a_118r.json
tail -5 $OUT/bench_r2a.err
