#!/bin/bash
# Weak scaling on N GPUs of one box: bash tools/gpu_scale.sh N tag
N=$1; TAG=$2; OUT=gpurun_out
nvidia-smi topo -m > $OUT/topo_$TAG.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 --profile-ranks $OUT/profile_ranks_$TAG.txt > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
d=json.loads(open('$OUT/bench_${TAG}_n$N.json').read().strip().splitlines()[-1])
print('N=$N case14: ms %.4f value %.3fM e2e %.3fM per-rank ms %s e2e GB/s %s'%(d['ms_per_step'],d['value']/1e6,d['e2e']['value']/1e6,d['config']['per_rank_ms_per_step'],d['config']['per_rank_e2e_d2h_gbs']))
for s in d.get('secondary',[]): print(s['workload'][:44],'value %.3fM e2e %.3fM ms %.3f per-rank %s'%(s['value']/1e6,s['e2e']/1e6,s['ms_per_step'],s['per_rank_ms_per_step']))
PY
tail -3 $OUT/bench_${TAG}_n$N.err
cat $OUT/profile_ranks_$TAG.txt
