"""Runs the UNMODIFIED reference (on oracle/shims) and oracle/flat.py side by side on the same environment folder and
action stream, and reports the largest deviation.  Build-container only (needs /root/reference).
    python tools/compare_oracle_reference.py <parameters_folder> <n_steps> <agent: nothing|random> [start_id] [mode]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
REF = os.environ.get('PYPOWNET_REFERENCE', '/root/reference')
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle', 'shims'), REF]


def reference_flag_code(flag):
    import pypownet.environment as E
    if flag is None:
        return 0
    if isinstance(flag, E.DivergingLoadflowException):
        return 2
    if isinstance(flag, E.TooManyConsumptionsCut):
        return 3
    if isinstance(flag, E.TooManyProductionsCut):
        return 4
    return 1


def random_action(rng, case, p_line=0.5, p_sub=0.5):
    """RandomNodeSplitting U RandomLineSwitch (agent.py:78-158), one substation + one line."""
    a = np.zeros(case.action_length, dtype=np.int64)
    if rng.random() < p_sub:
        s = rng.integers(case.n_sub)
        el = np.flatnonzero(case.elem_sub == s)
        a[el] = rng.integers(0, 2, size=len(el))
    if rng.random() < p_line:
        a[case.n_gen + case.n_load + 2 * case.n_line + rng.integers(case.n_line)] = 1
    return a


def main(folder, n_steps, agent, start_id=0, mode='soft', seed=0, verbose=True):
    import logging
    logging.disable(logging.CRITICAL)
    os.makedirs('/tmp/refrun_cwd', exist_ok=True)
    os.chdir('/tmp/refrun_cwd')
    from pypownet.environment import RunEnv
    from pypownet_b200.case import Case
    from pypownet_b200.chronic import ChronicSet
    from pypownet_b200.parameters import Parameters
    from oracle.flat import FlatEnv, Config
    t0 = time.time()
    env = RunEnv(folder, 'level0', start_id=start_id, game_over_mode=mode)
    par = Parameters(folder, 'level0')
    case = Case.from_file(par.get_reference_grid_path())
    chron = ChronicSet.from_folder(par.get_chronics_path())
    const = env.reward_signal.__dict__.get('too_many_productions_cut', None)
    cfg = Config(par.simulator_configuration, game_over_mode=mode, reward_constant=-const if const else None,
                 n_sub=case.n_sub)
    fe = FlatEnv(case, cfg, chron.chronics, start_id=start_id)
    o_ref = env._get_obs().as_array()
    o_fl = fe.observation()
    worst = float(np.max(np.abs(o_ref - o_fl)))
    rng = np.random.default_rng(seed)
    n_done = 0
    t_ref = t_fl = 0.0
    flags = {}
    for it in range(n_steps):
        a = np.zeros(case.action_length, dtype=np.int64) if agent == 'nothing' else random_action(rng, case)
        t1 = time.time()
        o1, r1, d1, f1 = env.step(a.copy(), do_sum=False)
        t2 = time.time()
        o2, r2, d2, f2, info = fe.step(a.copy())
        t3 = time.time()
        t_ref += t2 - t1
        t_fl += t3 - t2
        c1 = reference_flag_code(f1)
        flags[c1] = flags.get(c1, 0) + 1
        if d1 != d2 or c1 != f2 or (o1 is None) != (o2 is None):
            print('step %d: DISCRETE MISMATCH ref done=%s flag=%s (%s) | flat done=%s flag=%s' %
                  (it, d1, c1, getattr(f1, 'text', ''), d2, f2))
            return False
        dr = float(np.max(np.abs(np.asarray(r1, dtype=float) - r2))) if len(r1) == 5 else 0.0
        if o1 is not None:
            do = float(np.max(np.abs(o1[:len(o2)] - o2)))
            if do > 1e-6:
                k = int(np.argmax(np.abs(o1[:len(o2)] - o2)))
                print('step %d: obs deviation %g at index %d (ref %r flat %r)' % (it, do, k, o1[k], o2[k]))
                return False
            worst = max(worst, do)
        if dr > 1e-9:
            print('step %d: reward deviation %g ref %s flat %s flag %d' % (it, dr, r1, r2, c1))
            return False
        if d1:
            n_done += 1
            o1 = env.process_game_over()
            o2 = fe.process_game_over()
            do = float(np.max(np.abs(o1[:len(o2)] - o2)))
            if do > 1e-6:
                k = int(np.argmax(np.abs(o1[:len(o2)] - o2)))
                print('step %d: post-game-over obs deviation %g at index %d' % (it, do, k))
                return False
            worst = max(worst, do)
    if verbose:
        print('%s agent=%s steps=%d game-overs=%d flags=%s worst |obs diff|=%.3g  ref %.1f steps/s  flat %.1f steps/s'
              % (os.path.basename(folder), agent, n_steps, n_done, flags, worst, n_steps / t_ref, n_steps / t_fl))
    return True


if __name__ == '__main__':
    folder = os.path.abspath(sys.argv[1])
    ok = main(folder, int(sys.argv[2]), sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0,
              sys.argv[5] if len(sys.argv) > 5 else 'soft')
    sys.exit(0 if ok else 1)
