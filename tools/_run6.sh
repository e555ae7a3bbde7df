P='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("value %.4e ms %.4f n_gpus %d e2e %.4e"%(d["value"],d["ms_per_step"],d["n_gpus"],d["e2e"]["value"]))
    elif "rror" in l: print(l[:300])
'
echo "== 1 GPU case14"; timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu 2>&1 | python -c "$P"
echo "== 2 GPUs case14"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 2>&1 | python -c "$P"
