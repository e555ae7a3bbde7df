"""Loading of tests/golden/*.npz (made by tools/make_golden.py from the unmodified reference)."""
import glob
import json
import os

import numpy as np

from pypownet_b200.case import Case
from pypownet_b200.chronic import Chronic

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
TABLES = ('prods_p', 'prods_v', 'loads_p', 'loads_q', 'prods_p_planned', 'prods_v_planned', 'loads_p_planned',
          'loads_q_planned', 'maintenance', 'hazards')


def fixture_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, '*.npz')))


class Fixture(object):
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
        self.name = name
        self.z = z
        self.case = Case.builtin(str(z['casename']))
        self.config = json.loads(str(z['config']))
        self.mode = str(z['mode'])
        self.default_reward = bool(z['default_reward'])
        self.reward_constant = float(z['reward_constant'])
        self.thermal_limits = z['thermal_limits']
        self.chronics = []
        for i in range(int(z['n_chronics'])):
            # tables in the fixture are already shifted/truncated the way the reference holds them in memory
            self.chronics.append(_chronic_from_arrays(str(z['chronic%d_name' % i]),
                                                      {t: z['chronic%d_%s' % (i, t)] for t in TABLES},
                                                      z['chronic%d_ids' % i], z['chronic%d_datetimes' % i],
                                                      self.thermal_limits))
        for k in ('obs0', 'actions', 'obs', 'reward', 'done', 'flag', 'reset_obs'):
            setattr(self, k, z[k])
        # steps where the unmodified reference and the oracle disagree on done/flag (floating pockets: the reference's
        # outcome there is SuperLU rounding noise).  The reference's state after such a step is
        # recorded as state rows (resync_*) and every replay continues from it.
        n = len(self.actions)
        # 0: none; 1: done / flag of the step differ; 2: step agrees, only the restart (process_game_over) differs
        self.mismatch = z['mismatch'].astype(np.int8) if 'mismatch' in z.files else np.zeros(n, dtype=np.int8)
        self.resync = {}
        for k, t in enumerate(np.flatnonzero(self.mismatch)):
            self.resync[int(t)] = (z['resync_real'][k], z['resync_topo'][k], z['resync_cnt'][k])
        self.obs_width = self.obs.shape[1] if n else self.case.obs_length     # compact fixtures: dynamic prefix only
        self.has_sim = 'sim_actions' in z.files
        if self.has_sim:
            for k in ('sim_actions', 'sim_obs', 'sim_reward', 'sim_done', 'sim_flag'):
                setattr(self, k, z[k])
            self.sim_mismatch = z['sim_mismatch'] if 'sim_mismatch' in z.files else \
                np.zeros(len(self.sim_actions), dtype=bool)


def _chronic_from_arrays(name, tabs, ids, datetimes, imaps):
    ch = Chronic.__new__(Chronic)
    for t, a in tabs.items():
        setattr(ch, t, np.ascontiguousarray(a, dtype=np.float32))
    ch.name = name
    ch.ids = np.asarray(ids, dtype=np.int32)
    ch.datetimes = np.asarray(datetimes, dtype=np.int32).reshape(-1, 6)
    ch.imaps = np.asarray(imaps, dtype=np.float64)
    ch.n_rows = len(ch.ids)
    pos0 = np.flatnonzero(ch.ids == 0)
    ch.row_after_switch = min(int(pos0[0]) + 1, ch.n_rows - 1) if len(pos0) else -1
    return ch


def write_environment_folder(fx, root):
    """An environment folder in the reference's on-disk format (Appendix B of SURVEY.md) from a fixture, so that
    RunEnv(parameters_folder, 'level0') can be pointed at it.  Returns the parameters folder."""
    from oracle.ref_folder import write_environment_folder as write
    return write(root, fx.case, fx.config, fx.chronics, fx.reward_constant if fx.default_reward else None)
