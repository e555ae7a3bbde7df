"""ctypes binding of the C ABI in include/pypownet_b200.h (libpypownet_b200.so, built from csrc/ by
__graft_entry__.build()).  There is no fallback: without the CUDA library every compute entry point raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libpypownet_b200.so')

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)

FLAG_NONE, FLAG_ILLEGAL_ACTION, FLAG_DIVERGING_LOADFLOW, FLAG_TOO_MANY_LOADS_CUT, FLAG_TOO_MANY_PRODS_CUT = range(5)
STATE_REAL, STATE_TOPOLOGY, STATE_COUNTERS = 0, 1, 2


class PpnCase(C.Structure):
    _fields_ = [('n_sub', C.c_int32), ('n_gen', C.c_int32), ('n_load', C.c_int32), ('n_line', C.c_int32),
                ('base_mva', C.c_double),
                ('sub_ids', c_int32_p), ('gen_sub', c_int32_p), ('load_sub', c_int32_p),
                ('line_or_sub', c_int32_p), ('line_ex_sub', c_int32_p),
                ('line_r', c_double_p), ('line_x', c_double_p), ('line_b', c_double_p), ('line_tap', c_double_p),
                ('line_status0', c_uint8_p),
                ('bus_gs', c_double_p), ('bus_bs', c_double_p), ('bus_basekv', c_double_p),
                ('bus_vm0', c_double_p), ('bus_va0', c_double_p),
                ('gen_qmin', c_double_p), ('gen_qmax', c_double_p),
                ('gen_pg0', c_double_p), ('gen_qg0', c_double_p), ('gen_vg0', c_double_p),
                ('load_pd0', c_double_p), ('load_qd0', c_double_p),
                ('slack_sub', C.c_int32),
                ('thermal_limits', c_double_p)]


class PpnConfig(C.Structure):
    _fields_ = [('dc', C.c_int32),
                ('hard_overflow_coefficient', C.c_double),
                ('n_timesteps_hard_overflow_is_broken', C.c_int32),
                ('n_timesteps_consecutive_soft_overflow_breaks', C.c_double),
                ('n_timesteps_soft_overflow_is_broken', C.c_int32),
                ('n_timesteps_horizon_maintenance', C.c_int32),
                ('max_number_prods_game_over', C.c_int32),
                ('max_number_loads_game_over', C.c_int32),
                ('n_timesteps_actionned_line_reactionable', C.c_int32),
                ('n_timesteps_actionned_node_reactionable', C.c_int32),
                ('max_number_actionned_substations', C.c_int32),
                ('max_number_actionned_lines', C.c_int32),
                ('max_number_actionned_total', C.c_int32),
                ('hard_game_over', C.c_int32),
                ('loop_mode', C.c_int32),
                ('pf_tol', C.c_double),
                ('pf_max_it', C.c_int32),
                ('reward_constant', C.c_double),
                ('seed', C.c_uint64),
                ('threads_per_env', C.c_int32), ('pf_alg', C.c_int32)]


class PpnChronic(C.Structure):
    _fields_ = [('n_rows', C.c_int32),
                ('prods_p', c_float_p), ('prods_v', c_float_p), ('loads_p', c_float_p), ('loads_q', c_float_p),
                ('prods_p_planned', c_float_p), ('prods_v_planned', c_float_p),
                ('loads_p_planned', c_float_p), ('loads_q_planned', c_float_p),
                ('maintenance', c_float_p), ('hazards', c_float_p),
                ('ids', c_int32_p), ('datetimes', c_int32_p)]


# every symbol include/pypownet_b200.h declares: name -> (restype, argtypes)
VP = C.c_void_p
SYMBOLS = {
    'ppn_create': (C.c_int, [C.POINTER(PpnCase), C.POINTER(PpnConfig), C.c_int, C.c_int, C.POINTER(VP)]),
    'ppn_load_chronics': (C.c_int, [VP, C.c_int, C.POINTER(PpnChronic)]),
    'ppn_reset': (C.c_int, [VP, c_int32_p, c_int32_p, VP, C.c_int64, VP, VP]),
    'ppn_step': (C.c_int, [VP, VP, VP, C.c_int64, VP, VP, VP, VP, C.c_int, VP]),
    'ppn_simulate': (C.c_int, [VP, C.c_int, VP, VP, C.c_int64, VP, VP, VP, VP, VP]),
    'ppn_process_game_over': (C.c_int, [VP, VP, VP, C.c_int64, VP]),
    'ppn_action_valid': (C.c_int, [VP, VP, VP, VP]),
    'ppn_step_host': (C.c_int, [VP, VP, VP, C.c_int64, VP, VP, VP, VP, C.c_int]),
    'ppn_step_host_f32': (C.c_int, [VP, VP, VP, C.c_int64, VP, VP, VP, VP, C.c_int]),
    'ppn_set_result_pack': (C.c_int, [VP, VP]),
    'ppn_set_env_trace': (C.c_int, [VP, VP]),
    'ppn_sparse_selfcheck': (C.c_int, [C.c_int, C.c_int, VP, VP, C.c_int, C.c_uint32, VP, VP]),
    'ppn_state_width': (C.c_int, [VP, C.c_int]),
    'ppn_get_state': (C.c_int, [VP, C.c_int, VP, VP]),
    'ppn_set_state': (C.c_int, [VP, C.c_int, VP, VP]),
    'ppn_observation_static': (C.c_int, [VP, c_double_p]),
    'ppn_n_envs': (C.c_int, [VP]),
    'ppn_action_length': (C.c_int, [VP]),
    'ppn_obs_length': (C.c_int, [VP]),
    'ppn_obs_dynamic_length': (C.c_int, [VP]),
    'ppn_device': (C.c_int, [VP]),
    'ppn_get_counters': (C.c_int, [VP, C.POINTER(C.c_int64)]),
    'ppn_get_cascade_histogram': (C.c_int, [VP, C.POINTER(C.c_int64)]),
    'ppn_peer_alloc': (C.c_int, [C.c_int, C.c_uint64, C.POINTER(C.c_void_p), VP]),
    'ppn_peer_open': (C.c_int, [C.c_int, VP, C.POINTER(C.c_void_p)]),
    'ppn_peer_close': (C.c_int, [C.c_int, VP]),
    'ppn_peer_free': (C.c_int, [C.c_int, VP]),
    'ppn_peer_signal': (C.c_int, [C.c_int, VP, C.c_uint64, VP]),
    'ppn_peer_wait': (C.c_int, [C.c_int, VP, C.c_int, C.c_uint64, VP]),
    'ppn_peer_read': (C.c_int, [C.c_int, VP, VP, C.c_uint64, VP]),
    'ppn_last_error': (C.c_char_p, [VP]),
    'ppn_build_info': (C.c_char_p, []),
    'ppn_destroy': (None, [VP]),
}

_lib = None


class LibraryMissing(RuntimeError):
    pass


def load():
    """Load libpypownet_b200.so and type every entry point.  Raises LibraryMissing when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing('%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                             '(nvcc, sm_100a).  pypownet_b200 has no CPU path.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def as_ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def case_struct(case, thermal_limits):
    """PpnCase from a pypownet_b200.case.Case; returns (struct, keepalive list)."""
    keep = []

    def arr(x, dt):
        a = np.ascontiguousarray(x, dtype=dt)
        keep.append(a)
        return a

    S = case.n_sub
    loads_bus = case.load_sub.astype(np.int64)
    s = PpnCase()
    s.n_sub, s.n_gen, s.n_load, s.n_line = S, case.n_gen, case.n_load, case.n_line
    s.base_mva = case.base_mva
    s.sub_ids = as_ptr(arr(case.sub_ids, np.int32), C.c_int32)
    s.gen_sub = as_ptr(arr(case.gen_sub, np.int32), C.c_int32)
    s.load_sub = as_ptr(arr(case.load_sub, np.int32), C.c_int32)
    s.line_or_sub = as_ptr(arr(case.line_or_sub, np.int32), C.c_int32)
    s.line_ex_sub = as_ptr(arr(case.line_ex_sub, np.int32), C.c_int32)
    s.line_r = as_ptr(arr(case.line_r, np.float64), C.c_double)
    s.line_x = as_ptr(arr(case.line_x, np.float64), C.c_double)
    s.line_b = as_ptr(arr(case.line_b, np.float64), C.c_double)
    s.line_tap = as_ptr(arr(case.line_tap, np.float64), C.c_double)
    s.line_status0 = as_ptr(arr(case.line_status0, np.uint8), C.c_uint8)
    s.bus_gs = as_ptr(arr(case.bus_gs, np.float64), C.c_double)
    s.bus_bs = as_ptr(arr(case.bus_bs, np.float64), C.c_double)
    s.bus_basekv = as_ptr(arr(case.bus_basekv, np.float64), C.c_double)
    s.bus_vm0 = as_ptr(arr(case.bus_vm0, np.float64), C.c_double)
    s.bus_va0 = as_ptr(arr(case.bus_va0, np.float64), C.c_double)
    s.gen_qmin = as_ptr(arr(case.gen_qmin, np.float64), C.c_double)
    s.gen_qmax = as_ptr(arr(case.gen_qmax, np.float64), C.c_double)
    s.gen_pg0 = as_ptr(arr(case.gen_pg0, np.float64), C.c_double)
    s.gen_qg0 = as_ptr(arr(case.gen_qg0, np.float64), C.c_double)
    s.gen_vg0 = as_ptr(arr(case.gen_vg0, np.float64), C.c_double)
    s.load_pd0 = as_ptr(arr(case.bus_pd0[loads_bus], np.float64), C.c_double)
    s.load_qd0 = as_ptr(arr(case.bus_qd0[loads_bus], np.float64), C.c_double)
    s.slack_sub = case.slack_sub
    s.thermal_limits = as_ptr(arr(thermal_limits, np.float64), C.c_double)
    return s, keep


def config_struct(cfg, game_over_mode='soft', without_overflow_cutoff=False, loop_mode='natural', reward_constant=0.,
                  seed=0, threads_per_env=0):
    """PpnConfig from a configuration.yaml dict + the RunEnv arguments (game.py:263-298)."""
    s = PpnConfig()
    s.dc = 1 if str(cfg['loadflow_mode']).lower() == 'dc' else 0
    s.hard_overflow_coefficient = float(cfg['hard_overflow_coefficient'])
    s.n_timesteps_hard_overflow_is_broken = int(cfg['n_timesteps_hard_overflow_is_broken'])
    s.n_timesteps_consecutive_soft_overflow_breaks = float(cfg['n_timesteps_consecutive_soft_overflow_breaks'])
    s.n_timesteps_soft_overflow_is_broken = int(cfg['n_timesteps_soft_overflow_is_broken'])
    if without_overflow_cutoff:                                          # game.py:268-275
        s.hard_overflow_coefficient = 1e9
        s.n_timesteps_consecutive_soft_overflow_breaks = 1e12
    s.n_timesteps_horizon_maintenance = int(cfg['n_timesteps_horizon_maintenance'])
    s.max_number_prods_game_over = int(cfg['max_number_prods_game_over'])
    s.max_number_loads_game_over = int(cfg['max_number_loads_game_over'])
    s.n_timesteps_actionned_line_reactionable = int(cfg['n_timesteps_actionned_line_reactionable'])
    s.n_timesteps_actionned_node_reactionable = int(cfg['n_timesteps_actionned_node_reactionable'])
    s.max_number_actionned_substations = int(cfg['max_number_actionned_substations'])
    s.max_number_actionned_lines = int(cfg['max_number_actionned_lines'])
    s.max_number_actionned_total = int(cfg['max_number_actionned_total'])
    s.hard_game_over = 1 if game_over_mode == 'hard' else 0
    s.loop_mode = {'natural': 0, 'random': 1, 'fixed': 2}[loop_mode]
    s.pf_tol = 1e-6                                                      # grid.py:63
    s.pf_max_it = 25
    s.reward_constant = float(reward_constant)
    s.seed = int(seed)
    s.threads_per_env = int(threads_per_env)
    s.pf_alg = int(cfg.get('pf_alg', 2))          # 1: Newton-Raphson (not in the shipped configuration files)
    return s


def chronic_structs(chronics):
    """(ctypes array of PpnChronic, keepalive) from pypownet_b200.chronic.Chronic objects."""
    keep = []
    arr = (PpnChronic * len(chronics))()
    for i, ch in enumerate(chronics):
        s = arr[i]
        s.n_rows = ch.n_rows
        for name in ('prods_p', 'prods_v', 'loads_p', 'loads_q', 'prods_p_planned', 'prods_v_planned',
                     'loads_p_planned', 'loads_q_planned', 'maintenance', 'hazards'):
            a = np.ascontiguousarray(getattr(ch, name), dtype=np.float32)
            keep.append(a)
            setattr(s, name, as_ptr(a, C.c_float))
        ids = np.ascontiguousarray(ch.ids, dtype=np.int32)
        dts = np.ascontiguousarray(ch.datetimes, dtype=np.int32)
        keep += [ids, dts]
        s.ids = as_ptr(ids, C.c_int32)
        s.datetimes = as_ptr(dts, C.c_int32)
    return arr, keep
