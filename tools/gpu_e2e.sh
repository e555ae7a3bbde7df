#!/bin/bash
# e2e: zero-copy stores (default) against staged chunks + DMA, per grid
OUT=gpurun_out
run() { timeout 300 python bench.py --steps $4 --warmup 5 --no-cpu --no-secondary --grid $1 --envs $2 --agent $3 $5 > $OUT/tmp.json 2>> $OUT/bench_e2e.err
python - <<PY
import json
d=json.loads(open('$OUT/tmp.json').read().strip().splitlines()[-1])
print('$1 x $2 $3 $5 [$6]: kernel ms %.4f value %.3fM | e2e %.3fM (%.4f ms/step) e2e_f32 %.3fM'%(d['ms_per_step'],d['value']/1e6,d['e2e']['value']/1e6,1e3*$2/d['e2e']['value'],d['config']['e2e_float32_observations']['value']/1e6))
PY
}
run case118 8192 nothing 20 "" "zero-copy"
for c in 4 8 16; do PPN_HOST_STAGED=1 PPN_HOST_CHUNKS=$c run case118 8192 nothing 20 "" "staged $c chunks"; done
run case30 8192 nothing 30 --cascade "zero-copy"
for c in 4 8; do PPN_HOST_STAGED=1 PPN_HOST_CHUNKS=$c run case30 8192 nothing 30 --cascade "staged $c chunks"; done
run case14 4096 nothing 100 "" "zero-copy"
for c in 2 4; do PPN_HOST_STAGED=1 PPN_HOST_CHUNKS=$c run case14 4096 nothing 100 "" "staged $c chunks"; done
tail -3 $OUT/bench_e2e.err
