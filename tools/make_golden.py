"""Generates tests/golden/*.npz by running the UNMODIFIED reference package (/root/reference/pypownet) on the
oracle shims (oracle/shims: gym.spaces + restated PYPOWER).  Build-container only: /root/reference does not travel
to the GPU box, the fixtures do.

    python tools/make_golden.py [scenario ...]

Each fixture is self-contained: truncated chronic tables (float32, as parsed by the reference), configuration,
thermal limits, the action stream, and per step the reference's outputs through RunEnv.step / simulate /
process_game_over (environment.py:848-888): full observation vector, five sub-rewards, done, flag code.
oracle/flat.py is run in lockstep.  The FULL reference run is recorded: a step where reference and oracle disagree
on a discrete outcome (only grids with a floating pocket, whose outcome in the reference is SuperLU rounding noise --
oracle/flat.py header) is kept with `mismatch[t] = True`, and the state of the reference after
that step (after its process_game_over when it ended the game) is stored as state rows of the CUDA library
(`resync_*`), from which the oracle here and every replay in tests/ continue.  Nothing is truncated; tests report
the mismatch count.
"""
import json
import os
import shutil
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
REF = os.environ.get('PYPOWNET_REFERENCE', '/root/reference')
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle', 'shims'), REF]
OUT = os.path.join(ROOT, 'tests', 'golden')
TABLES = ('prods_p', 'prods_v', 'loads_p', 'loads_q', 'prods_p_planned', 'prods_v_planned', 'loads_p_planned',
          'loads_q_planned', 'maintenance', 'hazards')

# name: (source folder, builtin case, chronics kept, rows kept per chronic (None = all), n_steps, agent, mode,
#        config overrides, simulate each step, seed)
P = os.path.join(REF, 'parameters')
T = os.path.join(REF, 'tests', 'parameters')
SCENARIOS = {
    'd14_ac_nothing': (P + '/default14', 'case14', 'ab', 130, 300, 'nothing', 'soft', {}, False, 0),
    'd14_ac_random': (P + '/default14', 'case14', 'ab', 130, 220, 'random', 'soft', {}, True, 0),
    'd14_ac_random_hard': (P + '/default14', 'case14', 'abc', 60, 150, 'random', 'hard', {}, False, 3),
    'd14_dc_random': (P + '/default14', 'case14', 'ab', 130, 111, 'random', 'soft', {'loadflow_mode': 'DC'}, True, 7),
    'd14_dc_nothing': (P + '/default14', 'case14', 'ab', 130, 200, 'nothing', 'soft', {'loadflow_mode': 'DC'}, False, 0),
    'd14_tests_beta_dc': (T + '/default14_for_tests_beta', 'case14', 'a', None, 40, 'lines', 'soft', {}, False, 5),
    'd14_tests_hard_overflow': (T + '/default14_for_tests_hard_overflow', 'case14', 'a', None, 60, 'lines', 'soft',
                                {}, False, 2),
    'd14_tests_alpha': (T + '/default14_for_tests_alpha', 'case14', 'a', None, 60, 'lines', 'soft', {}, False, 4),
    'd14_tests_basic': (T + '/default14_for_tests', 'case14', 'a', None, 17, 'nothing', 'soft', {}, False, 0),
    'd30_ac_nothing': (P + '/default30', 'case30', 'ab', 100, 150, 'nothing', 'soft', {}, False, 0),
    'd30_ac_random': (P + '/default30', 'case30', 'ab', 100, 120, 'random', 'soft', {}, True, 5),
    'd118_ac_nothing': (P + '/default118', 'case118', 'ab', 40, 50, 'nothing', 'soft', {}, False, 0),
    'd118_ac_random': (P + '/default118', 'case118', 'ab', 40, 50, 'random', 'soft', {}, False, 6),
    # the shipped 600 A limits of default30 never trip; the synthetic limits of
    # tools/make_cascade_limits.py (pypownet_b200/data/case30.json, imaps_cascade) make the cascading-failure loop fire
    'd30_ac_cascade': (P + '/default30', 'case30', 'ab', 100, 150, 'nothing', 'soft', {'_imaps': 'case30'}, False, 0),
    'd118_ac_cascade': (P + '/default118', 'case118', 'ab', 40, 60, 'nothing', 'soft', {'_imaps': 'case118'}, False, 0),
    # BASELINE.json configs[0]: default14 DC, do-nothing agent, 1000 timesteps, single env (chronic a rolls into b)
    'd14_dc_nothing_1000': (P + '/default14', 'case14', 'ab', None, 1000, 'nothing', 'soft', {'loadflow_mode': 'DC'},
                            False, 0),
    # Newton-Raphson (PF_ALG = 1): the north star's named solver, NOT what the reference runs -- recorded by forcing the
    # shim's ppoption (PYPOWNET_SHIM_PF_ALG); `pf_alg` in the configuration selects it in the oracle and the library
    'd14_nr_random': (P + '/default14', 'case14', 'ab', 130, 160, 'random', 'soft', {'pf_alg': 1}, True, 11),
    'd30_nr_nothing': (P + '/default30', 'case30', 'ab', 100, 100, 'nothing', 'soft', {'pf_alg': 1}, False, 0),
    'd118_nr_nothing': (P + '/default118', 'case118', 'ab', 40, 40, 'nothing', 'soft', {'pf_alg': 1}, False, 0),
}
COMPACT = ('d14_dc_nothing_1000',)       # observations stored as their dynamic prefix only (static tail = obs0's)


def build_folder(src, chronics, rows, overrides, dst):
    """A truncated copy of an environment folder the reference can be pointed at."""
    import yaml
    if os.path.exists(dst):
        shutil.rmtree(dst)
    os.makedirs(os.path.join(dst, 'level0', 'chronics'))
    for f in os.listdir(src):
        if os.path.isfile(os.path.join(src, f)):
            shutil.copy(os.path.join(src, f), os.path.join(dst, f))
    for f in ('reference_grid.py', 'reference_grid.m'):
        shutil.copy(os.path.join(src, 'level0', f), os.path.join(dst, 'level0', f))
    with open(os.path.join(src, 'level0', 'configuration.yaml')) as f:
        cfg = yaml.safe_load(f)
    cfg.update({k: v for k, v in overrides.items() if not k.startswith('_')})
    imaps = overrides.get('_imaps')          # thermal limits that replace the chronics' own (the cascade scenarios)
    with open(os.path.join(dst, 'level0', 'configuration.yaml'), 'w') as f:
        yaml.safe_dump(cfg, f)
    for ch in chronics:
        s, d = os.path.join(src, 'level0', 'chronics', ch), os.path.join(dst, 'level0', 'chronics', ch)
        os.makedirs(d)
        for fn in os.listdir(s):
            with open(os.path.join(s, fn)) as f:
                lines = f.read().splitlines()
            keep = lines if (rows is None or fn == '_N_imaps.csv') else lines[:rows + 1]
            if fn == '_N_imaps.csv' and imaps is not None:
                assert len(imaps) == len(lines[0].split(';'))
                keep = [lines[0], ';'.join('%g' % v for v in imaps)]
            with open(os.path.join(d, fn), 'w') as f:
                f.write('\n'.join(keep) + '\n')
    return cfg


def flag_code(flag):
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


def make_action(rng, case, agent):
    """nothing | random (RandomNodeSplitting U RandomLineSwitch, agent.py:78-158) | lines (one random line switch
    every third step, to exercise cooldowns and reconnection of broken lines)."""
    a = np.zeros(case.action_length, dtype=np.uint8)
    if agent == 'random':
        if rng.random() < .5:
            s = rng.integers(case.n_sub)
            el = np.flatnonzero(case.elem_sub == s)
            a[el] = rng.integers(0, 2, size=len(el))
        if rng.random() < .5:
            a[case.n_gen + case.n_load + 2 * case.n_line + rng.integers(case.n_line)] = 1
        if rng.random() < .03:                              # sometimes far too many switches at once
            a[rng.integers(case.action_length, size=40)] = 1
    elif agent == 'lines':
        if rng.random() < .4:
            a[case.n_gen + case.n_load + 2 * case.n_line + rng.integers(case.n_line)] = 1
    return a


def reference_rows(game, case, chron):
    """State of the reference's Game (game.py:306-334, grid.py mpc tables) as the three state rows of the CUDA
    library (oracle.flat.FlatEnv.export_rows layout)."""
    grid = game.grid
    mpc = grid.mpc
    bus, gen, br = mpc['bus'], mpc['gen'], mpc['branch']
    S = case.n_sub
    real_ids = case.sub_ids.astype(np.int64)
    gnode = (gen[:, 0].astype(np.int64) != real_ids[case.gen_sub]).astype(np.uint8)
    onode = (br[:, 0].astype(np.int64) != real_ids[case.line_or_sub]).astype(np.uint8)
    enode = (br[:, 1].astype(np.int64) != real_ids[case.line_ex_sub]).astype(np.uint8)
    are_loads = np.asarray(grid.are_loads, dtype=bool)
    lnode = are_loads[case.load_sub + S].astype(np.uint8)
    assert np.all(are_loads[case.load_sub + S * lnode]) and are_loads.sum() == case.n_load
    lbus = case.load_sub + S * lnode
    real = np.concatenate((bus[:, 7], bus[:, 8], bus[lbus, 2], bus[lbus, 3], gen[:, 1], gen[:, 2], gen[:, 5]))
    topo = np.concatenate((gnode, lnode, onode, enode, (br[:, 10] != 0).astype(np.uint8),
                           (gen[:, 7] > 0).astype(np.uint8)))
    ch = game._Game__chronic
    names = [c.name for c in chron.chronics]
    ids = list(ch.get_timestep_ids())
    cnt = np.concatenate((game.timesteps_before_lines_reconnectable, game.timesteps_before_lines_reactionable,
                          game.n_timesteps_soft_overflowed_lines, game.timesteps_before_nodes_reactionable,
                          [names.index(ch.name), ids.index(game.current_timestep_id),
                           game._Game__chronic_looper.next_chronic_id, 0]))
    return real.astype(np.float64), topo.astype(np.uint8), cnt.astype(np.int32)


def run(name):
    import logging
    logging.disable(logging.CRITICAL)
    src, casename, chronics, rows, n_steps, agent, mode, overrides, do_sim, seed = SCENARIOS[name]
    if isinstance(overrides.get('_imaps'), str):
        with open(os.path.join(ROOT, 'pypownet_b200', 'data', overrides['_imaps'] + '.json')) as f:
            overrides = dict(overrides, _imaps=json.load(f)['imaps_cascade'])
    tmp = '/tmp/golden_envs/' + name
    cfgd = build_folder(src, chronics, rows, overrides, tmp)
    os.environ.pop('PYPOWNET_SHIM_PF_ALG', None)
    if overrides.get('pf_alg'):
        os.environ['PYPOWNET_SHIM_PF_ALG'] = str(overrides['pf_alg'])
    os.makedirs('/tmp/golden_cwd', exist_ok=True)
    os.chdir('/tmp/golden_cwd')
    from pypownet.environment import RunEnv
    from pypownet_b200.case import Case
    from pypownet_b200.chronic import ChronicSet
    from oracle.flat import FlatEnv, Config
    env = RunEnv(tmp, 'level0', game_over_mode=mode)
    case = Case.builtin(casename)
    assert np.array_equal(case.ppc['bus'], Case.from_file(os.path.join(tmp, 'level0', 'reference_grid.py')).ppc['bus'])
    chron = ChronicSet.from_folder(os.path.join(tmp, 'level0', 'chronics'))
    const = getattr(env.reward_signal, 'too_many_productions_cut', None)
    default_reward = const is not None
    cfg = Config(cfgd, game_over_mode=mode, reward_constant=-const if default_reward else 0., n_sub=case.n_sub)
    fe = FlatEnv(case, cfg, chron.chronics)
    rng = np.random.default_rng(seed)
    OBS = case.obs_length
    rec = {k: [] for k in ('actions', 'obs', 'reward', 'done', 'flag', 'reset_obs', 'sim_actions', 'sim_obs',
                           'sim_reward', 'sim_done', 'sim_flag', 'mismatch', 'sim_mismatch', 'resync_real',
                           'resync_topo', 'resync_cnt')}
    obs0 = env._get_obs().as_array()
    notes = []
    compact = name in COMPACT
    nd = case.obs_dynamic_length
    keep = (lambda o: o[:nd]) if compact else (lambda o: o)
    OW = nd if compact else OBS
    worst = float(np.max(np.abs(obs0 - fe.observation())))
    for it in range(n_steps):
        if do_sim:
            sa = make_action(rng, case, 'random')
            so, sr, sd, sf = env.simulate(sa.astype(np.int64), do_sum=False)
            so2, sr2, sd2, sf2, _ = fe.simulate(sa)
            sim_bad = sd != sd2 or flag_code(sf) != sf2
            if sim_bad:
                notes.append('simulate %d: reference done=%s flag=%d, oracle done=%s flag=%d' % (
                    it, sd, flag_code(sf), sd2, sf2))
            rec['sim_mismatch'].append(bool(sim_bad))
            rec['sim_actions'].append(sa)
            rec['sim_obs'].append(np.full(OW, np.nan) if so is None else keep(so))
            rec['sim_reward'].append(np.asarray(sr, dtype=np.float64) if len(sr) == 5 else np.full(5, np.nan))
            rec['sim_done'].append(bool(sd))
            rec['sim_flag'].append(flag_code(sf))
            if so is not None and not sim_bad:
                worst = max(worst, float(np.max(np.abs(so[:len(so2)] - so2))))
        a = make_action(rng, case, agent)
        o, r, d, f = env.step(a.astype(np.int64), do_sum=False)
        o2, r2, d2, f2, _ = fe.step(a)
        bad = d != d2 or flag_code(f) != f2
        if bad:
            notes.append('step %d: reference done=%s flag=%d (%s), oracle done=%s flag=%d' % (
                it, d, flag_code(f), getattr(f, 'text', ''), d2, f2))
        rec['mismatch'].append(1 if bad else 0)       # 1: done / flag of the step differ; 2 (below): only the restart differs
        rec['actions'].append(a)
        rec['obs'].append(np.full(OW, np.nan) if o is None else keep(o))
        rec['reward'].append(np.asarray(r, dtype=np.float64) if len(r) == 5 else np.full(5, np.nan))
        rec['done'].append(bool(d))
        rec['flag'].append(flag_code(f))
        if o is not None and not bad:
            worst = max(worst, float(np.max(np.abs(o[:len(o2)] - o2))))
        if d:
            try:
                ro = env.process_game_over()
            except RecursionError:
                # Newton-Raphson fixtures only: a NaN load-flow leaves NaN in gen QG, PYPOWER's makeSbus (sparse complex
                # product) turns that into NaN active injections, every later load-flow fails and the reference's
                # process_game_over recurses until Python gives up.  The fixture ends before this step.
                notes.append('stopped before step %d: the reference itself crashed (RecursionError in process_game_over)' % it)
                for k in ('mismatch', 'actions', 'obs', 'reward', 'done', 'flag'):
                    rec[k].pop()
                break
            rec['reset_obs'].append(keep(ro))
        else:
            rec['reset_obs'].append(np.full(OW, np.nan))
        if bad:
            # the replay continues from the reference's state (after its process_game_over, if it ended the game)
            rows = reference_rows(env.game, case, chron)
            fe.import_rows(*rows)
            for k, v in zip(('resync_real', 'resync_topo', 'resync_cnt'), rows):
                rec[k].append(v)
        elif d:
            ro2 = fe.process_game_over()
            dev = float(np.max(np.abs(ro[:len(ro2)] - ro2)))
            if not dev < 1e-6:
                # a floating pocket met inside process_game_over: the reference went on where the oracle restarted again
                notes.append('step %d: restart differs (max |reference - oracle| = %.3g)' % (it, dev))
                rec['mismatch'][-1] = 2
                rows = reference_rows(env.game, case, chron)
                fe.import_rows(*rows)
                for k, v in zip(('resync_real', 'resync_topo', 'resync_cnt'), rows):
                    rec[k].append(v)
            else:
                worst = max(worst, dev)
    n = len(rec['actions'])
    note = '; '.join(notes)
    out = {'casename': casename, 'config': json.dumps(cfgd), 'mode': mode, 'default_reward': default_reward,
           'reward_constant': cfg.reward_constant, 'thermal_limits': np.asarray(chron[0].imaps, dtype=np.float64),
           'obs0': obs0, 'note': note, 'n_chronics': len(chron), 'compact': compact}
    for i, ch in enumerate(chron.chronics):
        for t in TABLES:
            out['chronic%d_%s' % (i, t)] = getattr(ch, t)
        out['chronic%d_ids' % i] = ch.ids
        out['chronic%d_datetimes' % i] = ch.datetimes
        out['chronic%d_name' % i] = ch.name
    A = case.action_length
    out['actions'] = np.array(rec['actions'], dtype=np.uint8).reshape(n, A)
    out['obs'] = np.array(rec['obs'], dtype=np.float64).reshape(n, OW)
    out['reward'] = np.array(rec['reward'], dtype=np.float64).reshape(n, 5)
    out['done'] = np.array(rec['done'], dtype=bool)
    out['flag'] = np.array(rec['flag'], dtype=np.int32)
    out['reset_obs'] = np.array(rec['reset_obs'], dtype=np.float64).reshape(n, OW)
    out['mismatch'] = np.array(rec['mismatch'], dtype=np.int8)
    nm = int((out['mismatch'] != 0).sum())
    S_, G_, L_, N_ = case.n_sub, case.n_gen, case.n_load, case.n_line
    out['resync_real'] = np.array(rec['resync_real'], dtype=np.float64).reshape(nm, 4 * S_ + 2 * L_ + 3 * G_)
    out['resync_topo'] = np.array(rec['resync_topo'], dtype=np.uint8).reshape(nm, 2 * G_ + L_ + 3 * N_)
    out['resync_cnt'] = np.array(rec['resync_cnt'], dtype=np.int32).reshape(nm, 3 * N_ + S_ + 4)
    if do_sim:
        m = len(rec['sim_actions'])
        out['sim_actions'] = np.array(rec['sim_actions'], dtype=np.uint8).reshape(m, A)
        out['sim_obs'] = np.array(rec['sim_obs'], dtype=np.float64).reshape(m, OW)
        out['sim_reward'] = np.array(rec['sim_reward'], dtype=np.float64).reshape(m, 5)
        out['sim_done'] = np.array(rec['sim_done'], dtype=bool)
        out['sim_flag'] = np.array(rec['sim_flag'], dtype=np.int32)
        out['sim_mismatch'] = np.array(rec['sim_mismatch'], dtype=bool)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **out)
    flags = np.bincount(out['flag'], minlength=5).tolist()
    line = '%-26s steps=%4d/%4d game-overs=%3d flags[none,illegal,diverging,loads,prods]=%s ' \
           'max|reference-oracle|=%.2e pocket-mismatches=%d (+%d simulate) size=%dKB %s' % (
               name, n, n_steps, int(out['done'].sum()), flags, worst, nm,
               int(out['sim_mismatch'].sum()) if do_sim else 0, os.path.getsize(path) // 1024, note)
    print(line)
    return line


if __name__ == '__main__':
    names = sys.argv[1:] or list(SCENARIOS)
    lines = [run(n) for n in names]
    if not sys.argv[1:]:
        with open(os.path.join(OUT, 'MANIFEST.txt'), 'w') as f:
            f.write('Fixtures generated by tools/make_golden.py from the unmodified reference package on oracle/shims.\n'
                    'The reference\'s own suite on the same stack: 26 passed (pytest /root/reference/tests).\n\n')
            f.write('\n'.join(lines) + '\n')
