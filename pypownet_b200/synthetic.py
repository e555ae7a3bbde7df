"""Synthetic chronics with the statistics of the reference's shipped ones, generated from a grid case alone.

The shipped chronics (parameters/default14/level0/chronics/a..l, 75 MB of CSV) do not travel with this repo; the
benchmark configurations of BASELINE.json are defined on *synthetic* chronics of the same shape.  Measured on
default14/chronics/a (728 rows): active loads = 1.2 x case-file Pd on average with a daily cycle and 20 % spread,
reactive loads constant, set-point voltages constant, each generator switched off on 4-9 % of the rows,
maintenance on 1.9 % and hazards on 1.1 % of (row, line) pairs with durations 1 or 2, planned values within 5 % of
the realised ones, one row per hour.  Tables are float32 like the reference's parser (chronic.py:174-175).
"""
import numpy as np

from pypownet_b200.chronic import Chronic


def make_chronic(case, n_rows, seed, name=None, p_maintenance=0.019, p_hazard=0.011, p_gen_off=0.06,
                 thermal_limits=None, load_level=None):
    rng = np.random.default_rng(seed)
    G, L, N, S = case.n_gen, case.n_load, case.n_line, case.n_sub
    T = int(n_rows)
    t = np.arange(T + 1)                                   # one more row: `planned[t] := planned[t+1]`
    phase = rng.uniform(0, 24)
    daily = 1.0 + 0.17 * np.sin(2 * np.pi * (t - 9 + phase) / 24.) + 0.07 * np.sin(4 * np.pi * (t + phase) / 24.)
    if load_level is None:                                 # mean load / case-file load in the shipped chronics
        load_level = 1.2 if S <= 14 else (1.03 if S <= 30 else 0.9)
    level = load_level * daily * (1 + 0.03 * rng.standard_normal(T + 1))
    base_p = case.bus_pd0[case.load_sub]
    base_q = case.bus_qd0[case.load_sub]
    loads_p = base_p[None, :] * level[:, None] * (1 + 0.05 * rng.standard_normal((T + 1, L)))
    loads_q = np.repeat(base_q[None, :], T + 1, axis=0)
    # dispatch: on-line generators share 1.03 x the total load in proportion to fixed weights
    w = np.where(case.gen_pg0 > 0, case.gen_pg0, 0.)
    if w.sum() <= 0:
        w = np.ones(G)
    w = 0.55 * w / w.sum() + 0.45 / G
    on = rng.random((T + 1, G)) >= p_gen_off
    on[on.sum(axis=1) == 0, 0] = True
    share = w[None, :] * on
    share = share / share.sum(axis=1, keepdims=True)
    prods_p = share * (1.03 * loads_p.sum(axis=1))[:, None] * (1 + 0.05 * rng.standard_normal((T + 1, G)))
    gen_kv = case.bus_basekv[case.gen_sub]
    prods_v = np.where(on, (case.gen_vg0 * gen_kv)[None, :], 0.)
    prods_p = np.where(on, prods_p, 0.)

    def planned(x, keep_zero=None):
        y = x * (1 + 0.05 * rng.standard_normal(x.shape))
        return y if keep_zero is None else np.where(keep_zero, y, 0.)
    maintenance = (rng.random((T + 1, N)) < p_maintenance) * rng.integers(1, 3, size=(T + 1, N))
    hazards = (rng.random((T + 1, N)) < p_hazard) * rng.integers(1, 3, size=(T + 1, N))
    maintenance[0] = 0
    hazards[0] = 0
    hours = np.arange(T + 1)
    day = hours // 24
    datetimes = np.stack([np.full(T + 1, 2012), 1 + (day // 28) % 12, 1 + day % 28, hours % 24, np.zeros(T + 1),
                          np.zeros(T + 1)], axis=1).astype(np.int32)
    if thermal_limits is None:
        thermal_limits = np.full(N, 1e5)
    f32 = np.float32
    return Chronic(name or 'synthetic%d' % seed, prods_p.astype(f32), prods_v.astype(f32), loads_p.astype(f32),
                   loads_q.astype(f32), planned(prods_p, on).astype(f32), prods_v.astype(f32),
                   planned(loads_p).astype(f32), loads_q.astype(f32), maintenance.astype(f32), hazards.astype(f32),
                   np.arange(T + 1, dtype=np.int32), datetimes, np.asarray(thermal_limits, dtype=f32))


def make_chronics(case, n_chronics=12, n_rows=720, seed=0, thermal_limits=None, **kw):
    return [make_chronic(case, n_rows, seed * 1000 + i, name='syn%02d' % i, thermal_limits=thermal_limits, **kw)
            for i in range(n_chronics)]


# configuration.yaml of the shipped default environments (parameters/default{14,30,118}/level0/configuration.yaml)
DEFAULT_CONFIG = {
    'loadflow_backend': 'pypower', 'loadflow_mode': 'AC', 'max_seconds_per_timestep': 1.0,
    'hard_overflow_coefficient': 1.5, 'n_timesteps_hard_overflow_is_broken': 10,
    'n_timesteps_consecutive_soft_overflow_breaks': 3, 'n_timesteps_soft_overflow_is_broken': 5,
    'n_timesteps_horizon_maintenance': 20, 'max_number_prods_game_over': 1, 'max_number_loads_game_over': 0,
    'n_timesteps_actionned_line_reactionable': 3, 'n_timesteps_actionned_node_reactionable': 3,
    'n_timesteps_pending_line_reactionable_when_overflowed': 1,
    'n_timesteps_pending_node_reactionable_when_overflowed': 1,
    'max_number_actionned_substations': 7, 'max_number_actionned_lines': 10, 'max_number_actionned_total': 15,
}


def default_config(casename, **overrides):
    cfg = dict(DEFAULT_CONFIG)
    if casename == 'case30':
        cfg.update(max_number_actionned_substations=10, max_number_actionned_lines=15, max_number_actionned_total=20)
    elif casename == 'case118':
        cfg.update(hard_overflow_coefficient=2.0, n_timesteps_hard_overflow_is_broken=4,
                   n_timesteps_soft_overflow_is_broken=2, max_number_prods_game_over=10, max_number_loads_game_over=5,
                   max_number_actionned_substations=25, max_number_actionned_lines=40, max_number_actionned_total=50)
    cfg.update(overrides)
    return cfg
