"""VecRunEnv: B independent copies of one pypownet environment stepped together on one GPU.

Batched counterpart of pypownet.environment.RunEnv (environment.py:788-914): `step`, `simulate`,
`process_game_over`, `reset`, with the per-env results of the reference's tuple
`(observation | None, reward, done, flag)` returned as tensors with a leading env dimension.  All arithmetic is
done by the CUDA library behind the C ABI (include/pypownet_b200.h); PyTorch only provides device buffers and the
current stream.
"""
import ctypes as C

import numpy as np
import torch

from pypownet_b200 import _lib
from pypownet_b200.case import Case
from pypownet_b200.chronic import ChronicSet
from pypownet_b200.parameters import Parameters


class PpnError(RuntimeError):
    pass


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class VecRunEnv(object):
    def __init__(self, case, config, chronics, n_envs, device=0, game_over_mode='soft',
                 without_overflow_cutoff=False, loop_mode='natural', reward_constant=None, thermal_limits=None,
                 seed=0, threads_per_env=0, start_chronics=None, start_rows=None):
        """case: pypownet_b200.case.Case; config: dict of configuration.yaml; chronics: list of Chronic."""
        if not torch.cuda.is_available():
            raise PpnError('pypownet_b200 needs a CUDA device: the step path has no CPU implementation')
        self.lib = _lib.load()
        self.case, self.config, self.chronics = case, dict(config), list(chronics)
        self.n_envs = int(n_envs)
        self.device = torch.device('cuda', device)
        self.device_index = device
        if thermal_limits is None:
            # the limits of the chronic an env STARTS on, never refreshed (game.py:301-304); one set per handle: envs
            # that start on different chronics must agree (the shipped chronics of an environment all do)
            starts = [0] if start_chronics is None else sorted(set(int(c) for c in np.asarray(start_chronics).ravel()))
            thermal_limits = self.chronics[starts[0]].imaps
            for c in starts[1:]:
                if not np.array_equal(np.asarray(self.chronics[c].imaps), np.asarray(thermal_limits)):
                    raise ValueError('the envs of one handle start on chronics with different thermal limits '
                                     '(%s, %s): pass thermal_limits explicitly' % (self.chronics[starts[0]].name,
                                                                                   self.chronics[c].name))
        self.thermal_limits = np.asarray(thermal_limits, dtype=np.float64)
        if reward_constant is None:
            reward_constant = float(case.n_sub)
        cs, self._keep_case = _lib.case_struct(case, self.thermal_limits)
        cf = _lib.config_struct(self.config, game_over_mode, without_overflow_cutoff, loop_mode, reward_constant, seed,
                                threads_per_env)
        handle = C.c_void_p()
        self._check(self.lib.ppn_create(C.byref(cs), C.byref(cf), self.n_envs, device, C.byref(handle)), None)
        self.handle = handle
        arr, keep = _lib.chronic_structs(self.chronics)
        self._check(self.lib.ppn_load_chronics(self.handle, len(self.chronics), arr))
        del keep
        self.action_length = self.lib.ppn_action_length(self.handle)
        self.obs_length = self.lib.ppn_obs_length(self.handle)
        self.obs_dynamic_length = self.lib.ppn_obs_dynamic_length(self.handle)
        B = self.n_envs
        dev = self.device
        # the observation buffer holds full as_array rows; the static tail is written once here
        tail = np.zeros(self.obs_length - self.obs_dynamic_length, dtype=np.float64)
        self._check(self.lib.ppn_observation_static(self.handle, tail.ctypes.data_as(_lib.c_double_p)))
        self.obs_static = tail
        self.obs = torch.zeros((B, self.obs_length), dtype=torch.float64, device=dev)
        self.obs[:, self.obs_dynamic_length:] = torch.from_numpy(tail).to(dev)
        self.reward = torch.zeros((B, 5), dtype=torch.float64, device=dev)
        self.done = torch.zeros((B,), dtype=torch.uint8, device=dev)
        self.flag = torch.zeros((B,), dtype=torch.int32, device=dev)
        self.illegal = torch.zeros((B, 1 + 2 * case.n_line + case.n_sub), dtype=torch.uint8, device=dev)
        self.reset(start_chronics, start_rows)

    # ------------------------------------------------------------------------------------------------------------
    @classmethod
    def from_folder(cls, parameters_folder, game_level='level0', n_envs=1, chronic_looping_mode='natural', start_id=0,
                    game_over_mode='soft', without_overflow_cutoff=False, **kw):
        """Same arguments as RunEnv (environment.py:789-812) plus n_envs; every env starts on chronic start_id."""
        par = Parameters(parameters_folder, game_level)
        case = Case.from_file(par.get_reference_grid_path())
        chron = ChronicSet.from_folder(par.get_chronics_path())
        const = getattr(par.get_reward_signal_class()(), 'too_many_productions_cut', None) \
            if par.get_reward_signal_class() is not None else None
        kw.setdefault('reward_constant', -const if const is not None else float(case.n_sub))
        if not 0 <= int(start_id) < len(chron):          # the reference indexes its chronic list with start_id
            raise IndexError('start_id %d out of range: %d chronics in %s' % (start_id, len(chron), par.get_chronics_path()))
        kw.setdefault('start_chronics', np.full(n_envs, int(start_id), dtype=np.int32))
        env = cls(case, par.simulator_configuration, chron.chronics, n_envs, game_over_mode=game_over_mode,
                  without_overflow_cutoff=without_overflow_cutoff, loop_mode=chronic_looping_mode, **kw)
        env.parameters = par
        return env

    def _check(self, rc, handle='self'):
        if rc != 0:
            h = getattr(self, 'handle', None) if handle == 'self' else None
            msg = self.lib.ppn_last_error(h)
            raise PpnError('%s (code %d)' % (msg.decode() if msg else 'pypownet_b200 call failed', rc))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.ppn_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------------------
    def reset(self, start_chronics=None, start_rows=None):
        """Game.__init__ for every env: pristine grid, first row, first cascade.  Returns the observation tensor
        [B, obs_length] (a view of the env's buffer)."""
        B = self.n_envs
        ci = None if start_chronics is None else np.ascontiguousarray(start_chronics, dtype=np.int32)
        r0 = None if start_rows is None else np.ascontiguousarray(start_rows, dtype=np.int32)
        if ci is not None and ci.shape != (B,) or r0 is not None and r0.shape != (B,):
            raise ValueError('start_chronics / start_rows must have one entry per env')
        with torch.cuda.device(self.device):
            self._check(self.lib.ppn_reset(self.handle,
                                           None if ci is None else ci.ctypes.data_as(_lib.c_int32_p),
                                           None if r0 is None else r0.ctypes.data_as(_lib.c_int32_p),
                                           _ptr(self.obs), self.obs_length, _ptr(self.flag), self._stream()))
        return self.obs

    def _action_tensor(self, actions, rows):
        if actions is None:
            return None
        if isinstance(actions, np.ndarray):
            actions = torch.from_numpy(np.ascontiguousarray(actions, dtype=np.uint8))
        if actions.dtype != torch.uint8:
            actions = actions.to(torch.uint8)
        if tuple(actions.shape) != (rows, self.action_length):
            raise ValueError('Expected actions of shape (%d, %d), got %s' % (rows, self.action_length,
                                                                             tuple(actions.shape)))
        return actions.to(self.device, non_blocking=True).contiguous()

    def step(self, actions=None, auto_reset=False, want_obs=True):
        """RunEnv.step for every env.  actions: uint8 [B, action_length] (None = do-nothing).
        Returns (obs [B, obs_length], reward [B, 5], done [B] uint8, flag [B] int32) device tensors owned by the env
        (overwritten by the next call).  Rows of `obs` whose env is done keep their previous content unless
        auto_reset, in which case they hold the observation after process_game_over (Runner.step, runner.py:84-87)."""
        a = self._action_tensor(actions, self.n_envs)
        with torch.cuda.device(self.device):
            self._check(self.lib.ppn_step(self.handle, _ptr(a), _ptr(self.obs) if want_obs else None, self.obs_length,
                                          _ptr(self.reward), _ptr(self.done), _ptr(self.flag), _ptr(self.illegal),
                                          1 if auto_reset else 0, self._stream()))
        return self.obs, self.reward, self.done, self.flag

    def enable_result_pack(self, buffer=None):
        """From now on `step` also fills self.pack [B, 7] float64 = reward[5] | done | flag, written by the step kernel:
        the row an env-sharded run all-gathers each step (pypownet_b200.sharding.gather_results).  `buffer` switches
        to a caller-provided tensor (double buffering while a gather of the previous rows is still in flight)."""
        if buffer is not None:
            if buffer.dtype != torch.float64 or tuple(buffer.shape) != (self.n_envs, 7) or not buffer.is_contiguous():
                raise ValueError('result pack must be a contiguous float64 tensor of shape (%d, 7)' % self.n_envs)
            self.pack = buffer
        elif getattr(self, 'pack', None) is None:
            self.pack = torch.zeros((self.n_envs, 7), dtype=torch.float64, device=self.device)
        self._check(self.lib.ppn_set_result_pack(self.handle, _ptr(self.pack)))
        return self.pack

    def simulate(self, actions, n_candidates=1):
        """RunEnv.simulate for n_candidates actions per env ([B * n_candidates, action_length], env-major), no state
        change.  Returns fresh tensors (obs, reward, done, flag) with B * n_candidates rows."""
        rows = self.n_envs * n_candidates
        a = self._action_tensor(actions, rows)
        dev = self.device
        obs = torch.zeros((rows, self.obs_length), dtype=torch.float64, device=dev)
        obs[:, self.obs_dynamic_length:] = torch.from_numpy(self.obs_static).to(dev)
        reward = torch.zeros((rows, 5), dtype=torch.float64, device=dev)
        done = torch.zeros((rows,), dtype=torch.uint8, device=dev)
        flag = torch.zeros((rows,), dtype=torch.int32, device=dev)
        # illegality masks of the simulated actions (has_too_much_activations | reconnections | line cooldowns | substations)
        self.sim_illegal = torch.zeros((rows, 1 + 2 * self.case.n_line + self.case.n_sub), dtype=torch.uint8, device=dev)
        with torch.cuda.device(self.device):
            self._check(self.lib.ppn_simulate(self.handle, n_candidates, _ptr(a), _ptr(obs), self.obs_length,
                                              _ptr(reward), _ptr(done), _ptr(flag), _ptr(self.sim_illegal), self._stream()))
        return obs, reward, done, flag

    def process_game_over(self, mask=None):
        """RunEnv.process_game_over for the envs selected by mask (uint8/bool [B]; None = all)."""
        m = None
        if mask is not None:
            if isinstance(mask, np.ndarray):
                mask = torch.from_numpy(mask)
            m = mask.to(self.device).to(torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            self._check(self.lib.ppn_process_game_over(self.handle, _ptr(m), _ptr(self.obs), self.obs_length,
                                                       self._stream()))
        return self.obs

    def is_action_valid(self, actions):
        a = self._action_tensor(actions, self.n_envs)
        valid = torch.zeros((self.n_envs,), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.ppn_action_valid(self.handle, _ptr(a), _ptr(valid), self._stream()))
        return valid.bool()

    def step_host(self, actions, obs_out=None, auto_reset=False):
        """Host-buffer entry point (ppn_step_host): numpy in, numpy out, copies and synchronisation inside."""
        B = self.n_envs
        a = None if actions is None else np.ascontiguousarray(actions, dtype=np.uint8)
        if a is not None and a.shape != (B, self.action_length):
            raise ValueError('Expected actions of shape (%d, %d)' % (B, self.action_length))
        reward = np.empty((B, 5), dtype=np.float64)
        done = np.empty((B,), dtype=np.uint8)
        flag = np.empty((B,), dtype=np.int32)
        op = None
        stride = self.obs_dynamic_length
        if obs_out is not None:
            stride = obs_out.shape[1]
            op = obs_out.ctypes.data_as(C.c_void_p)
        self._check(self.lib.ppn_step_host(self.handle, None if a is None else a.ctypes.data_as(C.c_void_p), op, stride,
                                           reward.ctypes.data_as(C.c_void_p), done.ctypes.data_as(C.c_void_p),
                                           flag.ctypes.data_as(C.c_void_p), None, 1 if auto_reset else 0))
        return obs_out, reward, done, flag

    def step_pinned(self, actions_pinned, auto_reset=True, obs_dtype=torch.float64):
        """End-to-end step for a host-side agent through ppn_step_host: actions in pinned host memory (uint8
        [B, action_length] tensor, None = do-nothing) -> GPU, step, dynamic observation / reward / done / flag written by
        the step kernel straight into pinned host tensors (reused per call).  Synchronous.  obs_dtype=torch.float32 asks
        for float32 observation rows (ppn_step_host_f32: half the bytes over PCIe; the reference's dtype is float64)."""
        if obs_dtype not in (torch.float64, torch.float32):
            raise ValueError('obs_dtype must be torch.float64 or torch.float32')
        key = '_pin64' if obs_dtype == torch.float64 else '_pin32'
        if not hasattr(self, key):
            B = self.n_envs
            align = 2 if obs_dtype == torch.float64 else 4   # 16-byte aligned rows: one TMA bulk store per row
            stride = (self.obs_dynamic_length + align - 1) & ~(align - 1)
            full = torch.empty((B, stride), dtype=obs_dtype).pin_memory()
            setattr(self, key, (full, full[:, :self.obs_dynamic_length],
                                torch.empty((B, 5), dtype=torch.float64).pin_memory(),
                                torch.empty((B,), dtype=torch.uint8).pin_memory(),
                                torch.empty((B,), dtype=torch.int32).pin_memory()))
        full, po, pr, pd, pf = getattr(self, key)
        if actions_pinned is not None and (actions_pinned.dtype != torch.uint8 or not actions_pinned.is_contiguous()
                                           or tuple(actions_pinned.shape) != (self.n_envs, self.action_length)):
            raise ValueError('Expected a contiguous uint8 tensor of shape (%d, %d)' % (self.n_envs, self.action_length))
        fn = self.lib.ppn_step_host if obs_dtype == torch.float64 else self.lib.ppn_step_host_f32
        self._check(fn(self.handle, _ptr(actions_pinned), _ptr(full), full.shape[1], _ptr(pr), _ptr(pd), _ptr(pf), None,
                       1 if auto_reset else 0))
        return po, pr, pd, pf

    # ------------------------------------------------------------------------------------------------------------
    def get_state(self, field):
        w = self.lib.ppn_state_width(self.handle, field)
        dt = {_lib.STATE_REAL: torch.float64, _lib.STATE_TOPOLOGY: torch.uint8, _lib.STATE_COUNTERS: torch.int32}[field]
        out = torch.empty((self.n_envs, w), dtype=dt, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.ppn_get_state(self.handle, field, _ptr(out), self._stream()))
        return out

    def set_state(self, field, value):
        w = self.lib.ppn_state_width(self.handle, field)
        dt = {_lib.STATE_REAL: torch.float64, _lib.STATE_TOPOLOGY: torch.uint8, _lib.STATE_COUNTERS: torch.int32}[field]
        v = value.to(self.device).to(dt).contiguous()
        if tuple(v.shape) != (self.n_envs, w):
            raise ValueError('state field %d has shape (%d, %d)' % (field, self.n_envs, w))
        with torch.cuda.device(self.device):
            self._check(self.lib.ppn_set_state(self.handle, field, _ptr(v), self._stream()))

    def counters(self):
        """dict of cumulative device counters (load-flows, FD iterations, env steps, resets, max cascade depth) and
        launch information."""
        out = (C.c_int64 * 10)()
        self._check(self.lib.ppn_get_counters(self.handle, out))
        keys = ('loadflows', 'fd_iterations', 'env_steps', 'resets', 'max_cascade_depth', 'kernel_launches',
                'smem_bytes_per_env', 'threads_per_env', 'max_loadflows_one_env_step', 'max_fd_iterations_one_env_step')
        d = dict(zip(keys, [int(v) for v in out]))
        hist = (C.c_int64 * 8)()
        self._check(self.lib.ppn_get_cascade_histogram(self.handle, hist))
        d['cascade_depth_histogram'] = [int(v) for v in hist]
        return d
