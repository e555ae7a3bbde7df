"""Analysis only (CPU, numpy): would B'^-1 / B''^-1 obtained by UPDATING a stored inverse (Sherman-Morrison / bordering,
DESIGN.md section 7.1) change a discrete outcome of the step path?  An updated inverse differs from a freshly computed
one by rounding of the order 1e-13 relative.  This script replays the bench workload on the numpy restatement of the path
(oracle/flat.py, patched in memory: the two LU solves of fdpf become products with explicit inverses carrying relative
Gaussian noise) and counts the env-steps whose done / flag change and the load-flows whose iteration count changes.

    python tools/inverse_noise_study.py [grid] [envs] [steps]        -> profiles/r2n_factor_reuse.txt (last section)
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def patched_oracle():
    src = open(os.path.join(ROOT, 'oracle', 'flat.py')).read()

    def rep(old, new):
        nonlocal src
        assert old in src, old
        src = src.replace(old, new, 1)
    rep('''                    lupp = scipy.linalg.lu_factor(Bpp[np.ix_(pq, pq)], check_finite=False)''',
        '''                    lupp = scipy.linalg.lu_factor(Bpp[np.ix_(pq, pq)], check_finite=False)
                    if NOISE[0] > 0:
                        i1 = np.linalg.inv(Bp[np.ix_(pvpq, pvpq)])
                        i2 = np.linalg.inv(Bpp[np.ix_(pq, pq)])
                        i1 = i1 * (1 + NOISE[0] * RNG[0].standard_normal(i1.shape))
                        i2 = i2 * (1 + NOISE[0] * RNG[0].standard_normal(i2.shape))''')
    rep('Va[pvpq] = Va[pvpq] - scipy.linalg.lu_solve(lup, P, check_finite=False)',
        'Va[pvpq] = Va[pvpq] - (i1 @ P if NOISE[0] > 0 else scipy.linalg.lu_solve(lup, P, check_finite=False))')
    rep('Vm[pq] = Vm[pq] - scipy.linalg.lu_solve(lupp, Q, check_finite=False)',
        'Vm[pq] = Vm[pq] - (i2 @ Q if NOISE[0] > 0 else scipy.linalg.lu_solve(lupp, Q, check_finite=False))')
    rep('''                if self.cfg.pf_alg != 1:
                    self.last_iterations = i''',
        '''                if self.cfg.pf_alg != 1:
                    self.last_iterations = i
                    ITS.append((i, bool(success)))''')
    mod = types.ModuleType('flat_study')
    mod.__dict__.update({'NOISE': [0.0], 'RNG': [np.random.default_rng(0)], 'ITS': [], '__file__': os.path.join(ROOT, 'oracle', 'flat.py'),
                         '__package__': 'oracle'})
    sys.path.insert(0, ROOT)
    exec(compile(src, 'oracle/flat.py (patched)', 'exec'), mod.__dict__)
    return mod


def main():
    grid = sys.argv[1] if len(sys.argv) > 1 else 'case14'
    n_env = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    n_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    F = patched_oracle()
    case, cfg, chronics, imaps = bench.build_workload(grid, cascade=(grid == 'case30'))
    c, r = bench.shard_starts(4096, 0, 1)
    a = np.zeros(case.action_length, dtype=np.uint8)

    def run(noise):
        F.NOISE[0] = noise
        F.RNG[0] = np.random.default_rng(1)
        F.ITS.clear()
        out = []
        for env_id in range(n_env):
            k = (257 * env_id) % 4096
            fe = F.FlatEnv(case, F.Config(cfg, reward_constant=float(case.n_sub), n_sub=case.n_sub), chronics,
                           start_id=int(c[k]), thermal_limits=imaps, start_row=int(r[k]))
            for _ in range(n_steps):
                o, _, d, f = fe.step(a)[:4]
                out.append((bool(d), int(f), None if d else np.asarray(o)[:case.obs_dynamic_length].copy()))
                if d:
                    fe.process_game_over()
        return out, list(F.ITS)
    base, its0 = run(0.0)
    for noise in (1e-15, 1e-13, 1e-11, 1e-9):
        o, its = run(noise)
        same = sum(1 for x, y in zip(base, o) if x[0] == y[0] and x[1] == y[1])
        dif_it = sum(1 for x, y in zip(its0, its) if x != y) if len(its) == len(its0) else -1
        worst = max((float(np.max(np.abs(x[2] - y[2]))) for x, y in zip(base, o)
                     if x[2] is not None and y[2] is not None), default=0.0)
        print('%s: relative noise %.0e on both inverses: %d / %d env-steps with the same done and flag, %d load-flows, %d with '
              'another iteration count or outcome, max |observation difference| %.2e'
              % (grid, noise, same, len(base), len(its0), dif_it, worst))


if __name__ == '__main__':
    main()
