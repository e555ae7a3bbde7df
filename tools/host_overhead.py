"""Host-side cost of one VecRunEnv.step call (tiny batch, so the kernel itself is short) and kernel-only time at
several batch sizes.  Run on the GPU box."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pypownet_b200.vec_env import VecRunEnv  # noqa: E402

grid = sys.argv[1] if len(sys.argv) > 1 else 'case14'
tpe = int(sys.argv[3]) if len(sys.argv) > 3 else 0
case, cfg, chronics, imaps = bench.build_workload(grid)
for B in [int(x) for x in (sys.argv[2].split(',') if len(sys.argv) > 2 else '2,4096,16384,65536'.split(','))]:
    sc, sr = bench.env_starts(B)
    env = VecRunEnv(case, cfg, chronics, B, reward_constant=float(case.n_sub), thermal_limits=imaps,
                    start_chronics=sc, start_rows=sr, threads_per_env=tpe)
    act = torch.zeros((B, case.action_length), dtype=torch.uint8, device='cuda')
    for _ in range(20):
        env.step(act, auto_reset=True)
    torch.cuda.synchronize()
    n = 200
    t0 = time.perf_counter()
    for _ in range(n):
        env.step(act, auto_reset=True)
    t_launch = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        env.step(act, auto_reset=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    print('%s B=%6d  host enqueue %.1f us/step, wall %.1f us/step, device %.1f us/step -> %.3g env-steps/s  %s'
          % (grid, B, 1e6 * t_launch / n, 1e6 * t_all / n, 1e3 * ms, B / (ms * 1e-3), env.counters()))
    env.close()
