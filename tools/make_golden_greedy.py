"""Golden fixture for the tree-search agent: runs the UNMODIFIED reference GreedySearch (pypownet/agent.py:227-325) on
the unmodified reference environment (oracle shims underneath, as tools/make_golden.py) and records, per step, every
candidate action it simulated with the five sub-rewards the reference returned, the action it chose, and the outcome
of stepping with it.  Build-container only.

    python tools/make_golden_greedy.py          -> tests/golden/greedy/d14_greedy.npz
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import make_golden as mg  # noqa: E402  (sets sys.path for the reference and the shims)

N_STEPS = 8


def main():
    import logging
    logging.disable(logging.CRITICAL)
    src, casename, chronics, rows = mg.P + '/default14', 'case14', 'ab', 130
    tmp = '/tmp/golden_envs/d14_greedy'
    cfgd = mg.build_folder(src, chronics, rows, {}, tmp)
    os.makedirs('/tmp/golden_cwd', exist_ok=True)
    os.chdir('/tmp/golden_cwd')
    from pypownet.environment import RunEnv
    from pypownet.agent import GreedySearch
    from pypownet_b200.case import Case
    from pypownet_b200.chronic import ChronicSet
    env = RunEnv(tmp, 'level0', game_over_mode='soft')
    case = Case.builtin(casename)
    chron = ChronicSet.from_folder(os.path.join(tmp, 'level0', 'chronics'))
    const = env.reward_signal.too_many_productions_cut
    agent = GreedySearch(env)
    agent.verbose = False
    log = []
    real_simulate = env.simulate

    def recording_simulate(action, do_sum=True):
        out = real_simulate(action, do_sum=do_sum)
        a = action.as_array() if hasattr(action, 'as_array') else np.asarray(action)
        log.append((np.asarray(a, dtype=np.uint8), np.asarray(out[1], dtype=np.float64), bool(out[2]), mg.flag_code(out[3])))
        return out
    env.simulate = recording_simulate
    obs = env._get_obs().as_array()
    obs0 = obs.copy()
    rec = {k: [] for k in ('cand_actions', 'cand_reward', 'cand_done', 'cand_flag', 'actions', 'obs', 'reward', 'done',
                           'flag', 'reset_obs')}
    OBS, A = case.obs_length, case.action_length
    for it in range(N_STEPS):
        del log[:]
        action = agent.act(obs)
        a = np.asarray(action.as_array() if hasattr(action, 'as_array') else action, dtype=np.uint8)
        rec['cand_actions'].append(np.array([l[0] for l in log]))
        rec['cand_reward'].append(np.array([l[1] if len(l[1]) == 5 else np.full(5, np.nan) for l in log]))
        rec['cand_done'].append(np.array([l[2] for l in log]))
        rec['cand_flag'].append(np.array([l[3] for l in log], dtype=np.int32))
        o, r, d, f = env.step(action, do_sum=False)
        rec['actions'].append(a)
        rec['obs'].append(np.full(OBS, np.nan) if o is None else o)
        rec['reward'].append(np.asarray(r, dtype=np.float64))
        rec['done'].append(bool(d))
        rec['flag'].append(mg.flag_code(f))
        if d:
            o = env.process_game_over()
            rec['reset_obs'].append(o)
        else:
            rec['reset_obs'].append(np.full(OBS, np.nan))
        obs = o
        print('step %d: %d candidates, chose %s, best reward %.6f, done %s' % (
            it, len(log), np.flatnonzero(a).tolist(), np.nanmax(rec['cand_reward'][-1].sum(axis=1)), d))
    out = {'casename': casename, 'config': json.dumps(cfgd), 'mode': 'soft', 'default_reward': True,
           'reward_constant': -const, 'thermal_limits': np.asarray(chron[0].imaps, dtype=np.float64), 'obs0': obs0,
           'note': '', 'n_chronics': len(chron)}
    for i, ch in enumerate(chron.chronics):
        for t in mg.TABLES:
            out['chronic%d_%s' % (i, t)] = getattr(ch, t)
        out['chronic%d_ids' % i] = ch.ids
        out['chronic%d_datetimes' % i] = ch.datetimes
        out['chronic%d_name' % i] = ch.name
    n = len(rec['actions'])
    out['actions'] = np.array(rec['actions'], dtype=np.uint8).reshape(n, A)
    out['obs'] = np.array(rec['obs']).reshape(n, OBS)
    out['reward'] = np.array(rec['reward']).reshape(n, 5)
    out['done'] = np.array(rec['done'], dtype=bool)
    out['flag'] = np.array(rec['flag'], dtype=np.int32)
    out['reset_obs'] = np.array(rec['reset_obs']).reshape(n, OBS)
    out['cand_actions'] = np.array(rec['cand_actions'], dtype=np.uint8)
    out['cand_reward'] = np.array(rec['cand_reward'])
    out['cand_done'] = np.array(rec['cand_done'], dtype=bool)
    out['cand_flag'] = np.array(rec['cand_flag'], dtype=np.int32)
    os.makedirs(os.path.join(mg.OUT, 'greedy'), exist_ok=True)
    path = os.path.join(mg.OUT, 'greedy', 'd14_greedy.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KB;', out['cand_actions'].shape, 'candidates per step')


if __name__ == '__main__':
    main()
