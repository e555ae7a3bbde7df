#!/bin/bash
# Does the peer gather slow the IEEE-30 step at N=2?   (2 GPUs; PPN_BENCH_PG=none leaves the gather out)
OUT=gpurun_out
A="--grid case30 --envs 8192 --cascade --steps 60 --warmup 5 --no-cpu --no-secondary"
for d in "" none ""; do
PPN_BENCH_PG=$d timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 $A > $OUT/two_d.json 2>/dev/null
python - "$d" <<'PY'
import json, sys
try:
    d = json.loads(open('gpurun_out/two_d.json').read().strip().splitlines()[-1])
    print('case30 N=2, PPN_BENCH_PG=%-10s per rank kernel ms %s' % (sys.argv[1] or '(full)', d['config']['per_rank_ms_per_step']))
except Exception as ex:
    print('failed', ex)
PY
done
