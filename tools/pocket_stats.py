"""How often does the reference meet a grid with a floating pocket (a multi-bus component without the reference
bus), and what does it return then?  Runs the UNMODIFIED reference (on oracle/shims) with a random node-splitting /
line-switching agent and inspects every load-flow.  Build-container only.
    python tools/pocket_stats.py <parameters_folder> <n_steps> [DC]"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
REF = os.environ.get('PYPOWNET_REFERENCE', '/root/reference')
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle', 'shims'), REF, os.path.join(ROOT, 'tools')]


def main(folder, n_steps, dc=False, agent='random'):
    import logging
    import warnings
    logging.disable(logging.CRITICAL)
    warnings.simplefilter('ignore')
    import make_golden as mg
    tmp = '/tmp/pocket_env_%s_%s' % (os.path.basename(folder), 'dc' if dc else 'ac')
    mg.build_folder(folder, 'abc', None, {'loadflow_mode': 'DC'} if dc else {}, tmp)
    os.makedirs('/tmp/golden_cwd', exist_ok=True)
    os.chdir('/tmp/golden_cwd')
    from pypownet.environment import RunEnv
    import pypownet.grid as G
    from pypownet_b200.case import Case
    stats = {'loadflows': 0, 'pocket': 0, 'pocket_ok': 0, 'pocket_diverged': 0, 'nopocket_diverged': 0}
    orig = G.Grid.compute_loadflow

    def has_pocket(mpc):
        br = mpc['branch']
        on = br[:, 10] != 0
        ids = np.unique(np.r_[br[on, 0], br[on, 1]])
        idx = {b: i for i, b in enumerate(ids)}
        lab = np.arange(len(ids))
        f = np.array([idx[b] for b in br[on, 0]], dtype=int)
        t = np.array([idx[b] for b in br[on, 1]], dtype=int)
        while True:
            new = lab.copy()
            np.minimum.at(new, f, lab[t])
            np.minimum.at(new, t, lab[f])
            if (new == lab).all():
                break
            lab = new
        return len(np.unique(lab)) > 1

    def patched(self, *a, **k):
        pocket = has_pocket(self.mpc)
        stats['loadflows'] += 1
        stats['pocket'] += pocket
        try:
            r = orig(self, *a, **k)
            stats['pocket_ok'] += pocket
            return r
        except G.DivergingLoadflowException:
            stats['pocket_diverged' if pocket else 'nopocket_diverged'] += 1
            raise
    G.Grid.compute_loadflow = patched
    env = RunEnv(tmp, 'level0')
    case = Case.from_file(os.path.join(tmp, 'level0', 'reference_grid.py'))
    rng = np.random.default_rng(0)
    for it in range(n_steps):
        a = mg.make_action(rng, case, agent)
        o, r, d, f = env.step(a.astype(np.int64), do_sum=False)
        if d:
            env.process_game_over()
    G.Grid.compute_loadflow = orig
    print('%s %s agent=%s steps=%d: %s' % (os.path.basename(folder), 'DC' if dc else 'AC', agent, n_steps, stats))
    return stats


if __name__ == '__main__':
    main(os.path.abspath(sys.argv[1]), int(sys.argv[2]), len(sys.argv) > 3 and sys.argv[3] == 'DC',
         sys.argv[4] if len(sys.argv) > 4 else 'random')
