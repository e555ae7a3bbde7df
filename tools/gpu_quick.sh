#!/bin/bash
# parity + the four bench workloads, one line each:  bash tools/gpu_quick.sh [label]
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/quick_tests.log 2>&1; tail -3 $OUT/quick_tests.log
run() { timeout 300 python bench.py --steps $4 --warmup 5 --no-cpu --no-secondary --grid $1 --envs $2 --agent $3 $5 > $OUT/tmp.json 2>> $OUT/bench_quick.err
python - <<PY
import json
d=json.loads(open('$OUT/tmp.json').read().strip().splitlines()[-1])
print('$1 x $2 $3 $5 [$6]: kernel ms %.4f value %.3fM | e2e %.3fM e2e_f32 %.3fM'%(d['ms_per_step'],d['value']/1e6,d['e2e']['value']/1e6,d['config']['e2e_float32_observations']['value']/1e6))
PY
}
run case14 4096 nothing 100 "" "$1"
run case14 4096 random 100 "" "$1"
run case30 8192 nothing 40 --cascade "$1"
run case118 8192 nothing 40 "" "$1"
run case118 4096 random 40 "" "$1"
tail -3 $OUT/bench_quick.err
