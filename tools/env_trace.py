"""Where does a step's time go?  Per-env trace of the bench workload (ppn_set_env_trace): SM cycles, load-flows,
fast-decoupled iterations and restarts of every env in every step.  A launch lasts as long as its slowest env.
    python tools/env_trace.py [grid] [envs] [steps] [random]   (GPU box)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pypownet_b200.vec_env import VecRunEnv, _ptr  # noqa: E402

grid = sys.argv[1] if len(sys.argv) > 1 else 'case14'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
agent = sys.argv[4] if len(sys.argv) > 4 else 'nothing'
case, cfg, chronics, imaps = bench.build_workload(grid)
sc, sr = bench.env_starts(B)
env = VecRunEnv(case, cfg, chronics, B, device=0, reward_constant=float(case.n_sub), thermal_limits=imaps,
                start_chronics=sc, start_rows=sr)
trace = torch.zeros((B, 4), dtype=torch.int64, device='cuda')
env._check(env.lib.ppn_set_env_trace(env.handle, _ptr(trace)))
bank = torch.from_numpy(bench.random_action_bank(case, B)).cuda() if agent == 'random' else None
for t in range(10):
    env.step(None if bank is None else bank[t % 16], auto_reset=True)
rows = []
X, Y = [], []
for t in range(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    env.step(None if bank is None else bank[t % 16], auto_reset=True)
    e1.record()
    torch.cuda.synchronize()
    tr = trace.cpu().numpy()
    cyc = tr[:, 0]
    top = np.argsort(-cyc)[:3]
    rows.append((e0.elapsed_time(e1), cyc.mean(), np.percentile(cyc, 50), np.percentile(cyc, 99), cyc.max()))
    print('step %2d: %.3f ms | cycles mean %6.0fk p50 %6.0fk p99 %6.0fk max %6.0fk | slowest envs (kcycles, load-flows, '
          'iterations, restarts): %s' % (t, rows[-1][0], rows[-1][1] / 1e3, rows[-1][2] / 1e3, rows[-1][3] / 1e3,
                                         rows[-1][4] / 1e3, [(int(cyc[i] // 1000), int(tr[i, 1]), int(tr[i, 2]),
                                                              int(tr[i, 3])) for i in top]))
    X.append(np.c_[tr[:, 1], tr[:, 2], np.ones(B)])
    Y.append(cyc)
X, Y = np.vstack(X), np.concatenate(Y)
coef = np.linalg.lstsq(X, Y, rcond=None)[0]
r = np.array(rows)
print('fit: cycles = %.0f per load-flow + %.0f per iteration + %.0f' % tuple(coef))
print('mean over steps: %.3f ms per step; env cycles mean %.0fk, max %.0fk (max / mean = %.1f)'
      % (r[:, 0].mean(), r[:, 1].mean() / 1e3, r[:, 4].mean() / 1e3, r[:, 4].mean() / r[:, 1].mean()))
