#!/bin/bash
# Round 2, visit d (TWO GPUs): result rows written into rank 0's memory over NVLink -- parity test + scaling at N=2.
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_peer_gather.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --profile-ranks $OUT/profile_ranks_r2d.txt > $OUT/bench_r2d_n2.json 2> $OUT/bench_r2d_n2.err; echo "bench n2 rc=$?"; tail -1 $OUT/bench_r2d_n2.json | cut -c1-2500; tail -5 $OUT/bench_r2d_n2.err
timeout 300 python bench.py --gpus 1 --steps 100 --warmup 10 --no-cpu --profile-ranks $OUT/profile_ranks_r2d.txt > $OUT/bench_r2d_n1.json 2> $OUT/bench_r2d_n1.err; tail -1 $OUT/bench_r2d_n1.json | cut -c1-600
cat $OUT/profile_ranks_r2d.txt
