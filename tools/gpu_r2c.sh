#!/bin/bash
# Round 2, visit c (one GPU): per-env trace of the bench workload, new bench.py with the secondary configurations.
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python tools/env_trace.py case14 4096 20 > $OUT/trace_r2c_14.txt 2>&1; tail -8 $OUT/trace_r2c_14.txt
timeout 300 python tools/env_trace.py case118 8192 6 > $OUT/trace_r2c_118.txt 2>&1; tail -4 $OUT/trace_r2c_118.txt
timeout 300 python tools/env_trace.py case118 4096 6 random > $OUT/trace_r2c_118r.txt 2>&1; tail -4 $OUT/trace_r2c_118r.txt
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > $OUT/bench_r2c.json 2> $OUT/bench_r2c.err; echo "bench rc=$?"; tail -1 $OUT/bench_r2c.json | cut -c1-3000; tail -5 $OUT/bench_r2c.err
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
