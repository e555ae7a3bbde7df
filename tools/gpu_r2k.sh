#!/bin/bash
# Round 2, visit k (one GPU): hybrid solve with single-warp sparse levels (IEEE-118), smaller shared-memory plan (IEEE-30).
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { timeout 300 python bench.py --steps $4 --warmup 5 --no-cpu --no-secondary --grid $1 --envs $2 --agent $3 $5 > $OUT/tmp.json 2>> $OUT/bench_r2k.err
python - <<PY
import json
d=json.loads(open('$OUT/tmp.json').read().strip().splitlines()[-1])
print('$1 x $2 $3 $5 $6: ms %.4f value %.3fM e2e %.3fM e2e_f32 %.3fM smem/env %d'%(d['ms_per_step'],d['value']/1e6,d['e2e']['value']/1e6,d['config']['e2e_float32_observations']['value']/1e6,d['config']['smem_bytes_per_env']))
PY
}
run case118 8192 nothing 30
run case118 4096 random 30
run case30 8192 nothing 50 --cascade
PPN_WARP_TABLES_SMEM=1 run case30 8192 nothing 50 --cascade "(tables staged per env, as before)"
PPN_EPB=1 run case30 8192 nothing 50 --cascade "(one env per CTA)"
run case30 8192 random 50 --cascade
run case14 4096 nothing 100
tail -3 $OUT/bench_r2k.err
