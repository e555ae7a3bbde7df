#!/bin/bash
# drain kernel of ppn_step_host: parity of the host step, then e2e with and without it per grid
OUT=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "host or pinned or float32 or step" > $OUT/drain_tests.log 2>&1; tail -3 $OUT/drain_tests.log
run() { timeout 300 python bench.py --steps $4 --warmup 5 --no-cpu --no-secondary --grid $1 --envs $2 --agent $3 $5 > $OUT/tmp.json 2>> $OUT/bench_drain.err
python - <<PY
import json
d=json.loads(open('$OUT/tmp.json').read().strip().splitlines()[-1])
print('$1 x $2 $3 $5 [$6]: kernel ms %.4f value %.3fM | e2e %.3fM (%.4f ms/step) e2e_f32 %.3fM'%(d['ms_per_step'],d['value']/1e6,d['e2e']['value']/1e6,1e3*$2/d['e2e']['value'],d['config']['e2e_float32_observations']['value']/1e6))
PY
}
PPN_HOST_DRAIN=0 run case14 4096 nothing 100 "" "no drain"
run case14 4096 nothing 100 "" "drain after the step kernel, priority stream"
run case14 4096 random 100 "" "drain after the step kernel, priority stream"
PPN_DRAIN_ROWS=32 run case14 4096 nothing 100 "" "drain 32 rows/warp"
PPN_DRAIN_ROWS=128 run case14 4096 nothing 100 "" "drain 128 rows/warp"
CUDA_LAUNCH_BLOCKING=1 run case14 4096 nothing 30 "" "kernels serialised (CUDA_LAUNCH_BLOCKING=1)"
tail -3 $OUT/bench_drain.err
