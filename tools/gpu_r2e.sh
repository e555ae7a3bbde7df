#!/bin/bash
# Round 2, visit e (TWO GPUs): block vs strided sharding at N=2, and the workload effect emulated on one GPU.
OUT=gpurun_out
for mode in blocks strided; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --sharding $mode --no-secondary --profile-ranks $OUT/profile_ranks_r2e.txt > $OUT/bench_r2e_n2_$mode.json 2> $OUT/bench_r2e.err; echo "bench n2 $mode rc=$?"; tail -1 $OUT/bench_r2e_n2_$mode.json | cut -c1-400
done
for sh in 0/1 0/2 1/2 3/8 7/8; do
for mode in blocks strided; do
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --no-secondary --sharding $mode --emulate-shard $sh > $OUT/tmp.json 2>> $OUT/bench_r2e.err
python - <<PY
import json
d=json.loads(open('$OUT/tmp.json').read().strip().splitlines()[-1])
print('emulate shard $sh $mode: ms %.4f value %.3fM e2e %.3fM'%(d['ms_per_step'],d['value']/1e6,d['e2e']['value']/1e6))
PY
done; done
cat $OUT/profile_ranks_r2e.txt; tail -3 $OUT/bench_r2e.err
