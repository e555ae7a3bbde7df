#!/bin/bash
# Round 2, one GPU: profiles of what ships.  bash tools/gpu_profiles.sh A|B   (two visits: at most 64 MiB come back per visit)
OUT=gpurun_out; TAG=${TAG:-r2n}
mkdir -p $OUT
prof() {  # tag grid envs agent extra
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ppn_step_kernel -s 8 -c 1 -f -o $OUT/prof_$1 python bench.py --grid $2 --envs $3 --agent $4 $5 --steps 8 --warmup 3 --no-cpu --no-secondary > $OUT/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
}
if [ "$1" = "A" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}14.csv python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/bench_under_ncu_${TAG}.log 2>&1; echo "launch list rc=$?"
  prof ${TAG}14 case14 4096 nothing
  prof ${TAG}30 case30 8192 nothing --cascade
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "float32 or fixture" 2>&1 | tail -3
  timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; tail -1 $OUT/bench_${TAG}.json | cut -c1-2200
else
  prof ${TAG}118 case118 8192 nothing
  prof ${TAG}118r case118 4096 random
  timeout 300 python tools/phase_timing.py case118 8192 > $OUT/phase_${TAG}_118.txt 2>&1; tail -12 $OUT/phase_${TAG}_118.txt
  timeout 300 python tools/phase_timing.py case14 4096 > $OUT/phase_${TAG}_14.txt 2>&1; tail -6 $OUT/phase_${TAG}_14.txt
fi
ls -la $OUT/*.ncu-rep
