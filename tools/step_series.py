"""Per-step device time series (events around each step kernel, no flush) to see how the cost evolves."""
import os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import bench
from pypownet_b200.vec_env import VecRunEnv
grid = sys.argv[1] if len(sys.argv) > 1 else 'case14'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
n = int(sys.argv[3]) if len(sys.argv) > 3 else 600
case, cfg, chronics, imaps = bench.build_workload(grid)
sc, sr = bench.env_starts(B)
env = VecRunEnv(case, cfg, chronics, B, reward_constant=float(case.n_sub), thermal_limits=imaps, start_chronics=sc, start_rows=sr)
act = torch.zeros((B, case.action_length), dtype=torch.uint8, device='cuda')
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
dones = []
c_prev = env.counters()
ev[0].record()
for k in range(n):
    o, r, d, f = env.step(act, auto_reset=True)
    ev[k + 1].record()
    dones.append(d.sum())
    if k % 50 == 49:
        c = env.counters(); print('after step', k, 'max LF per env-step', c['max_loadflows_one_env_step'], 'max FD iterations', c['max_fd_iterations_one_env_step'])
torch.cuda.synchronize()
t = np.array([ev[k].elapsed_time(ev[k + 1]) * 1e3 for k in range(n)])
dn = torch.stack(dones).cpu().numpy()
for a in range(0, n, 50):
    print('steps %3d-%3d: mean %.0f us  min %.0f max %.0f   game-overs/step %.1f' % (a, a + 49, t[a:a + 50].mean(), t[a:a + 50].min(), t[a:a + 50].max(), dn[a:a + 50].mean()))
print(env.counters())
