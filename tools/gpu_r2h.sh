#!/bin/bash
# Round 2, visit h (one GPU): Newton-Raphson option -- fixtures recorded from the reference forced to PF_ALG=1, ragged batches.
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python - <<'PY'
import sys, time, torch
sys.path.insert(0,'.')
import bench
from pypownet_b200.vec_env import VecRunEnv
for grid,B in (('case14',4096),('case30',4096),('case118',512)):
    case,cfg,chronics,imaps=bench.build_workload(grid)
    cfg=dict(cfg,pf_alg=1)
    sc,sr=bench.shard_starts(B,0,1)
    env=VecRunEnv(case,cfg,chronics,B,device=0,reward_constant=float(case.n_sub),thermal_limits=imaps,start_chronics=sc,start_rows=sr)
    for _ in range(3): env.step(None,auto_reset=True)
    torch.cuda.synchronize(); t0=time.perf_counter()
    n=10
    for _ in range(n): env.step(None,auto_reset=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t0
    c=env.counters()
    print('NR %s x %d: %.3f ms/step, %.3f M env-steps/s, %.2f iterations per load-flow'%(grid,B,1e3*dt/n,B*n/dt/1e6,c['fd_iterations']/max(c['loadflows'],1)))
PY
