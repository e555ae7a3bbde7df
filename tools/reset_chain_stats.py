"""How long are the restart chains of the bench workload?  Oracle run (CPU): per game over, the number of load-flows
process_game_over needs (1 = the first restart succeeds).  Motivates the speculative parallel restarts of DESIGN.md 7.
    python tools/reset_chain_stats.py [grid] [envs] [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle.flat import FlatEnv, Config  # noqa: E402

grid = sys.argv[1] if len(sys.argv) > 1 else 'case14'
n_envs = int(sys.argv[2]) if len(sys.argv) > 2 else 96
n_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 150
case, cfg, chronics, imaps = bench.build_workload(grid)
sc, sr = bench.env_starts(n_envs)
a = np.zeros(case.action_length, dtype=np.uint8)
chains, step_lf = [], []
for e in range(n_envs):
    env = FlatEnv(case, Config(cfg, reward_constant=float(case.n_sub), n_sub=case.n_sub), chronics, start_id=int(sc[e]),
                  thermal_limits=imaps, start_row=int(sr[e]))
    for t in range(n_steps):
        n0 = env.n_loadflows
        done = env.step(a)[2]
        n1 = env.n_loadflows
        if done:
            env.process_game_over()
            chains.append(env.n_loadflows - n1)
        step_lf.append(env.n_loadflows - n0)
chains, step_lf = np.array(chains), np.array(step_lf)
print('%s: %d env-steps, %d game overs (%.1f %%)' % (grid, len(step_lf), len(chains), 100.0 * len(chains) / len(step_lf)))
print('load-flows per env-step: mean %.2f, max %d' % (step_lf.mean(), step_lf.max()))
print('load-flows per restart: ' + ', '.join('%d: %.1f %%' % (k, 100.0 * np.mean(chains == k)) for k in range(1, 9))
      + ', more: %.1f %%' % (100.0 * np.mean(chains > 8)) + ', max %d' % chains.max())
