#!/bin/bash
# Round 2, visit j (one GPU): profiles of what ships.  bash tools/gpu_r2j.sh A|B   (two visits: at most 64 MiB come back per visit)
OUT=gpurun_out
mkdir -p $OUT
prof() {  # tag grid envs agent extra
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ppn_step_kernel -s 8 -c 1 -f -o $OUT/prof_$1 python bench.py --grid $2 --envs $3 --agent $4 $5 --steps 8 --warmup 3 --no-cpu --no-secondary > $OUT/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
}
if [ "$1" = "A" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_r2j14.csv python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/bench_under_ncu_r2j.log 2>&1; echo "launch list rc=$?"
  prof r2j14 case14 4096 nothing
  prof r2j30 case30 8192 nothing --cascade
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "float32 or fixture" 2>&1 | tail -3
  timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > $OUT/bench_r2j.json 2> $OUT/bench_r2j.err; tail -1 $OUT/bench_r2j.json | cut -c1-2200
else
  prof r2j118 case118 8192 nothing
  prof r2j118r case118 4096 random
  timeout 300 python tools/phase_timing.py case118 8192 > $OUT/phase_r2j_118.txt 2>&1; tail -12 $OUT/phase_r2j_118.txt
  timeout 300 python tools/phase_timing.py case14 4096 > $OUT/phase_r2j_14.txt 2>&1; tail -6 $OUT/phase_r2j_14.txt
fi
ls -la $OUT/*.ncu-rep
