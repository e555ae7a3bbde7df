#!/bin/bash
# Round 2, second visit: deferred restarts (K attempts side by side) -- parity, then the K sweep.
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/gputests_r2b.log 2>&1; echo "tests rc=$?"; tail -5 $OUT/gputests_r2b.log
for K in 0 1 2 4 6; do
  PPN_RESET_K=$K timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu > $OUT/bench_r2b_14_k$K.json 2>> $OUT/bench_r2b.err
  python - <<PY
import json
d=json.load(open('$OUT/bench_r2b_14_k$K.json'))
print('case14 K=$K ms %.4f value %.3fM warm %.3fM e2e %.3fM launches %d'%(d['ms_per_step'],d['value']/1e6,d['config']['warm_l2_value']/1e6,d['e2e']['value']/1e6,d['gpu_launches']))
PY
done
for K in 0 1 2; do
  PPN_RESET_K=$K timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --grid case118 --envs 8192 > $OUT/bench_r2b_118_k$K.json 2>> $OUT/bench_r2b.err
  PPN_RESET_K=$K timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --grid case30 --envs 8192 > $OUT/bench_r2b_30_k$K.json 2>> $OUT/bench_r2b.err
  python - <<PY
import json
for g in ('118','30'):
    d=json.load(open('$OUT/bench_r2b_%s_k$K.json'%g))
    print('case%s K=$K ms %.4f value %.3fM e2e %.3fM'%(g,d['ms_per_step'],d['value']/1e6,d['e2e']['value']/1e6))
PY
done
PPN_RESET_K=4 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --envs 16384 > $OUT/bench_r2b_14_16k_k4.json 2>> $OUT/bench_r2b.err
PPN_RESET_K=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --envs 16384 > $OUT/bench_r2b_14_16k_k0.json 2>> $OUT/bench_r2b.err
python - <<PY
import json
for k in (0,4):
    d=json.load(open('$OUT/bench_r2b_14_16k_k%d.json'%k))
    print('case14 16384 K=%d ms %.4f value %.3fM'%(k,d['ms_per_step'],d['value']/1e6))
PY
tail -5 $OUT/bench_r2b.err
