#!/bin/bash
# Round 2, visit f (one GPU): paired 16-byte loads in the mismatch gather / solve; spread starts.
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu > $OUT/bench_r2f.json 2> $OUT/bench_r2f.err
python - <<PY
import json
d=json.loads(open('$OUT/bench_r2f.json').read().strip().splitlines()[-1])
print('case14 spread: ms %.4f value %.3fM warm %.3fM e2e %.3fM frac %.4f'%(d['ms_per_step'],d['value']/1e6,d['config']['warm_l2_value']/1e6,d['e2e']['value']/1e6,d['roofline']['frac']))
for s in d['secondary']: print(s['workload'][:40],'value %.3fM e2e %.3fM ms %.3f'%(s['value']/1e6,s['e2e']/1e6,s['ms_per_step']))
PY
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --no-secondary --sharding blocks > $OUT/tmp.json 2>> $OUT/bench_r2f.err
python - <<PY
import json
d=json.loads(open('$OUT/tmp.json').read().strip().splitlines()[-1])
print('case14 blocks (round-1 workload): ms %.4f value %.3fM e2e %.3fM'%(d['ms_per_step'],d['value']/1e6,d['e2e']['value']/1e6))
PY
timeout 300 python tools/env_trace.py case14 4096 8 2>&1 | tail -4
tail -3 $OUT/bench_r2f.err
