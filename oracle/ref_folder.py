"""TEST INFRASTRUCTURE ONLY (oracle/): writes an environment folder in the reference's on-disk format (SURVEY.md
Appendix B: level0/{configuration.yaml, reference_grid.py, chronics/<name>/*.csv}) from in-memory tables, so that the
UNMODIFIED reference's RunEnv(parameters_folder, 'level0') can be pointed at a fixture or at the synthetic bench
workload.  Used by tests/golden_util.py and by bench.py's CPU-baseline legs; never by the product package."""
import datetime
import os

import numpy as np

_NAMES = {'prods_p': '_N_prods_p.csv', 'prods_v': '_N_prods_v.csv', 'loads_p': '_N_loads_p.csv',
          'loads_q': '_N_loads_q.csv', 'prods_p_planned': '_N_prods_p_planned.csv',
          'prods_v_planned': '_N_prods_v_planned.csv', 'loads_p_planned': '_N_loads_p_planned.csv',
          'loads_q_planned': '_N_loads_q_planned.csv', 'maintenance': 'maintenance.csv', 'hazards': 'hazards.csv'}


def write_environment_folder(root, case, config, chronics, reward_constant=None):
    """case: pypownet_b200.case.Case; config: dict of configuration.yaml; chronics: list of Chronic (planned tables
    already shifted the way the reference holds them in memory: the shift is undone here).  reward_constant: write a
    reward_signal.py with the shipped five-term reward and this constant.  Returns root."""
    import yaml

    from pypownet_b200.case import write_case_file
    level = os.path.join(root, 'level0')
    os.makedirs(os.path.join(level, 'chronics'), exist_ok=True)
    write_case_file(os.path.join(level, 'reference_grid.py'), case.ppc)
    with open(os.path.join(level, 'reference_grid.m'), 'w') as f:      # mandatory for parameters.py:35-40, only read by the
        f.write('% MATPOWER twin of reference_grid.py: not used with loadflow_backend pypower\n')   # Octave backend
    with open(os.path.join(level, 'configuration.yaml'), 'w') as f:
        yaml.safe_dump(dict(config), f)
    if reward_constant is not None:
        with open(os.path.join(root, 'reward_signal.py'), 'w') as f:
            f.write('from pypownet_b200.reward_signal import DefaultRewardSignal\n\n\n'
                    'class CustomRewardSignal(DefaultRewardSignal):\n'
                    '    def __init__(self):\n        super().__init__(constant=%r)\n' % reward_constant)
    for ch in chronics:
        d = os.path.join(level, 'chronics', ch.name)
        os.makedirs(d, exist_ok=True)
        for t, fn in _NAMES.items():
            a = np.asarray(getattr(ch, t), dtype=np.float32)
            if t.endswith('_planned'):                     # undo `planned[t] := planned[t+1]` (chronic.py:202-205)
                a = np.vstack([a[:1], a[:-1]])
            with open(os.path.join(d, fn), 'w') as f:
                f.write(';'.join('c%d' % i for i in range(a.shape[1])) + '\n')
                for row in a:
                    f.write(';'.join(repr(float(v)) for v in row) + '\n')
        with open(os.path.join(d, '_N_simu_ids.csv'), 'w') as f:
            f.write('simu_id\n' + '\n'.join('%.1f' % i for i in ch.ids) + '\n')
        with open(os.path.join(d, '_N_imaps.csv'), 'w') as f:
            f.write(';'.join('c%d' % i for i in range(len(ch.imaps))) + '\n' +
                    ';'.join(repr(float(v)) for v in ch.imaps) + '\n')
        with open(os.path.join(d, '_N_datetimes.csv'), 'w') as f:
            f.write('date;time\n')
            for y, mo, dd, h, mi, s in ch.datetimes:
                f.write(datetime.datetime(int(y), int(mo), int(dd), int(h), int(mi)).strftime('%Y-%b-%d;%H:%M').lower() + '\n')
    return root
