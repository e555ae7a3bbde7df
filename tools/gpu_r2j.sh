#!/bin/bash
# Round 2, visit j (one GPU): profiles of what ships -- launch list of the default bench command, one full ncu capture of
# the step kernel per BASELINE configuration, phase timings.
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_r2j14.csv python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/bench_under_ncu_r2j.log 2>&1; echo "launch list rc=$?"
prof() {  # tag grid envs agent extra
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ppn_step_kernel -s 8 -c 1 -f -o $OUT/prof_$1 python bench.py --grid $2 --envs $3 --agent $4 $5 --steps 8 --warmup 3 --no-cpu --no-secondary > $OUT/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
}
prof r2j14 case14 4096 nothing
prof r2j30 case30 8192 nothing --cascade
prof r2j118 case118 8192 nothing
prof r2j118r case118 4096 random
timeout 300 python tools/phase_timing.py case118 8192 > $OUT/phase_r2j_118.txt 2>&1; tail -12 $OUT/phase_r2j_118.txt
timeout 300 python tools/phase_timing.py case14 4096 > $OUT/phase_r2j_14.txt 2>&1; tail -6 $OUT/phase_r2j_14.txt
ls -la $OUT/*.ncu-rep
