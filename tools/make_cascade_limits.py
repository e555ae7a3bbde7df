"""Synthetic thermal limits that make the cascading-failure loop (game.py:503-589) fire on IEEE-30 / IEEE-118.

The shipped limits of default30 (600 A) and default118 (2000 A) are never exceeded, so a "default30 AC with
cascading-failure loop enabled" run (BASELINE.json configs[2]) never trips a line.  SURVEY.md 8(d) config 3 prescribes
limits' = 1.05 x the 90th percentile over time of each line's do-nothing Ampere flow.  This script computes them with
the CPU oracle (oracle/flat.py, shipped limits, bench chronics of pypownet_b200/synthetic.py, seed 0) and stores them
as `imaps_cascade` in pypownet_b200/data/<grid>.json.  Deterministic: no random draws beyond the seeded chronics.

    python tools/make_cascade_limits.py [case30 case118]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)


def limits(grid, steps_per_chronic=120):
    import bench
    from oracle.flat import FlatEnv, Config
    case, cfg, chronics, imaps = bench.build_workload(grid, cascade=False)
    G, L, N = case.n_gen, case.n_load, case.n_line
    off = 4 * L + 4 * G + 2 * N
    samples = []
    a = np.zeros(case.action_length, dtype=np.uint8)
    for c in range(len(chronics)):
        env = FlatEnv(case, Config(cfg, reward_constant=float(case.n_sub), n_sub=case.n_sub), chronics, start_id=c,
                      thermal_limits=imaps, start_row=17 * c)
        for _ in range(steps_per_chronic):
            o, r, d, f, _ = env.step(a)
            if d:
                o = env.process_game_over()
            amp, status = o[off:off + N], o[off + N:off + 2 * N]
            samples.append(np.where(status > 0, amp, np.nan))
    s = np.array(samples)
    p90 = np.nanpercentile(s, 90, axis=0)
    # whole amperes, like the shipped _N_imaps.csv files; at least 1 A
    return np.maximum(np.ceil(1.05 * p90), 1.0)


if __name__ == '__main__':
    for grid in (sys.argv[1:] or ['case30', 'case118']):
        lim = limits(grid)
        path = os.path.join(ROOT, 'pypownet_b200', 'data', grid + '.json')
        with open(path) as f:
            d = json.load(f)
        d['imaps_cascade'] = [float(v) for v in lim]
        with open(path, 'w') as f:
            json.dump(d, f)
        print(grid, 'imaps_cascade: min %.0f median %.0f max %.0f A' % (lim.min(), np.median(lim), lim.max()))
