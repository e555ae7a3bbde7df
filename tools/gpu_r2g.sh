#!/bin/bash
# Round 2, visit g (one GPU): bordered hybrid factor for split buses on IEEE-118 (AC).
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-secondary --grid case118 --envs 4096 --agent random > $OUT/bench_r2g_118r.json 2> $OUT/bench_r2g.err
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-secondary --grid case118 --envs 8192 > $OUT/bench_r2g_118.json 2>> $OUT/bench_r2g.err
python - <<PY
import json
for f in ('118r','118'):
    d=json.loads(open('$OUT/bench_r2g_%s.json'%f).read().strip().splitlines()[-1])
    print(f,'ms %.4f value %.3fM e2e %.3fM lf/step %.3f'%(d['ms_per_step'],d['value']/1e6,d['e2e']['value']/1e6,d['config']['loadflows_per_env_step']))
PY
timeout 300 python tools/env_trace.py case118 4096 4 random 2>&1 | tail -3
tail -3 $OUT/bench_r2g.err
