"""GPU parity: the CUDA step path (through the C ABI) against (1) the fixtures recorded from the unmodified
reference and (2) the CPU oracle on batches of envs started at different chronics / rows.
Tolerance: 1e-6 on every observation entry (BASELINE.json north_star: bus voltage magnitudes/angles to 1e-6), the
suite asserts the tighter 1e-7 seen in practice; line status, done and flag codes bit-exact."""
import numpy as np
import pytest
import torch

from golden_util import Fixture, fixture_names
from oracle.flat import FlatEnv, Config

pytestmark = pytest.mark.gpu
TOL = 1e-7


def vec_env(fx, B, **kw):
    from pypownet_b200.vec_env import VecRunEnv
    return VecRunEnv(fx.case, fx.config, fx.chronics, B, device=0, game_over_mode=fx.mode,
                     reward_constant=fx.reward_constant, thermal_limits=fx.thermal_limits, **kw)


def set_rows(env, rows, B):
    from pypownet_b200 import _lib
    for field, row in zip((_lib.STATE_REAL, _lib.STATE_TOPOLOGY, _lib.STATE_COUNTERS), rows):
        env.set_state(field, torch.from_numpy(np.repeat(np.asarray(row)[None], B, axis=0)))


@pytest.mark.parametrize('name', fixture_names())
def test_cuda_reproduces_reference_fixture(name):
    fx = Fixture(name)
    B = 3
    env = vec_env(fx, B)
    nd = fx.case.obs_dynamic_length
    W = fx.obs_width
    obs0 = env.obs.cpu().numpy()
    assert np.max(np.abs(obs0[0] - fx.obs0)) < TOL
    worst = 0.0
    for t in range(len(fx.actions)):
        if fx.mismatch[t] == 1:
            # floating pocket: the reference's outcome is decided by the rounding of a singular
            # SuperLU pivot; the library reports "diverging", and the replay continues from the reference's state
            obs, reward, done, flag = env.step(np.repeat(fx.actions[t][None], B, axis=0))
            assert np.all(done.cpu().numpy() == 1) and np.all(flag.cpu().numpy() == 2), t
            set_rows(env, fx.resync[t], B)
            continue
        if fx.has_sim and not fx.sim_mismatch[t]:
            so, sr, sd, sf = env.simulate(np.repeat(fx.sim_actions[t][None], B, axis=0))
            assert bool(sd[1].item()) == bool(fx.sim_done[t]) and int(sf[1].item()) == int(fx.sim_flag[t]), t
            if not fx.sim_done[t]:
                worst = max(worst, float(np.max(np.abs(so[1].cpu().numpy() - fx.sim_obs[t]))))
            if fx.default_reward:
                assert np.max(np.abs(sr[1].cpu().numpy() - fx.sim_reward[t])) < TOL
        obs, reward, done, flag = env.step(np.repeat(fx.actions[t][None], B, axis=0))
        d, f = done.cpu().numpy(), flag.cpu().numpy()
        assert np.all(d == int(fx.done[t])) and np.all(f == int(fx.flag[t])), 'step %d: %s %s' % (t, d, f)
        if fx.default_reward:
            assert np.max(np.abs(reward.cpu().numpy() - fx.reward[t][None])) < TOL, t
        if fx.mismatch[t] == 2:     # same outcome, but a floating pocket inside the reference's process_game_over
            set_rows(env, fx.resync[t], B)
            continue
        if fx.done[t]:
            obs = env.process_game_over(done)
            expect = fx.reset_obs[t]
        else:
            expect = fx.obs[t]
        got = obs.cpu().numpy()
        assert np.array_equal(got[0], got[B - 1])
        err = float(np.max(np.abs(got[0, :W] - expect)))
        assert err < TOL, 'step %d: |cuda - reference| = %g at %d' % (t, err, int(np.argmax(np.abs(got[0, :W] - expect))))
        worst = max(worst, err)
    assert worst < TOL
    print('%s: %d steps replayed, %d floating-pocket mismatches, max |cuda - reference| = %.3g'
          % (name, len(fx.actions), int((fx.mismatch != 0).sum()), worst))


@pytest.mark.parametrize('name,steps', [('d14_ac_random', 120), ('d30_ac_random', 60), ('d118_ac_random', 25),
                                        ('d14_dc_random', 60), ('d118_ac_random:DC', 20), ('d30_ac_random:DC', 30),
                                        ('d14_ac_random:NR', 80), ('d30_ac_random:NR', 40), ('d118_ac_random:NR', 12)])
def test_cuda_matches_oracle_on_a_ragged_batch(name, steps):
    """Envs start on different chronics and rows and receive different random actions; auto-reset as Runner does.
    `:DC` replays a grid's fixture data with loadflow_mode DC (rundcpf through the same sparse / hybrid solvers), `:NR`
    with the Newton-Raphson solver (the oracle's newtonpf is itself pinned on tests/golden/d*_nr_*.npz)."""
    name, _, mode = name.partition(':')
    fx = Fixture(name)
    if mode == 'DC':
        fx.config = dict(fx.config, loadflow_mode='DC')
    if mode == 'NR':                                  # Newton-Raphson (ppn_config.pf_alg = 1) against the oracle's newtonpf
        fx.config = dict(fx.config, pf_alg=1)
    B = 24
    rng = np.random.default_rng(11)
    nch = len(fx.chronics)
    start_c = rng.integers(0, nch, size=B).astype(np.int32)
    start_r = np.array([rng.integers(0, fx.chronics[c].n_rows - 1) for c in start_c], dtype=np.int32)
    start_r[0] = 0
    env = vec_env(fx, B, start_chronics=start_c, start_rows=start_r)
    cfg = Config(fx.config, game_over_mode=fx.mode, reward_constant=fx.reward_constant, n_sub=fx.case.n_sub)
    refs = [FlatEnv(fx.case, cfg, fx.chronics, start_id=int(start_c[e]), thermal_limits=fx.thermal_limits,
                    start_row=int(start_r[e])) for e in range(B)]
    nd = fx.case.obs_dynamic_length
    got0 = env.obs.cpu().numpy()
    for e in range(B):
        assert np.max(np.abs(got0[e, :nd] - refs[e].observation_dynamic())) < TOL, e
    case = fx.case
    worst = 0.0
    for t in range(steps):
        acts = np.zeros((B, case.action_length), dtype=np.uint8)
        for e in range(B):
            if rng.random() < .5:
                s = rng.integers(case.n_sub)
                el = np.flatnonzero(case.elem_sub == s)
                acts[e, el] = rng.integers(0, 2, size=len(el))
            if rng.random() < .5:
                acts[e, case.n_gen + case.n_load + 2 * case.n_line + rng.integers(case.n_line)] = 1
        obs, reward, done, flag = env.step(acts, auto_reset=True)
        got, r, d, f = obs.cpu().numpy(), reward.cpu().numpy(), done.cpu().numpy(), flag.cpu().numpy()
        for e in range(B):
            o2, r2, d2, f2, _ = refs[e].step(acts[e])
            assert (bool(d[e]), int(f[e])) == (bool(d2), int(f2)), 'step %d env %d' % (t, e)
            assert np.max(np.abs(r[e] - r2)) < TOL
            if d2:
                o2 = refs[e].process_game_over()
            err = float(np.max(np.abs(got[e, :nd] - o2)))
            assert err < TOL, 'step %d env %d: %g' % (t, e, err)
            worst = max(worst, err)
    c = env.counters()
    assert c['env_steps'] >= B * steps and c['loadflows'] >= c['env_steps']


def test_step_host_matches_device_entry_point():
    fx = Fixture('d14_ac_nothing')
    B = 5
    e1, e2 = vec_env(fx, B), vec_env(fx, B)
    obs_h = np.zeros((B, fx.case.obs_dynamic_length))
    for t in range(30):
        a = np.repeat(fx.actions[t][None], B, axis=0)
        o1, r1, d1, f1 = e1.step(a, auto_reset=True)
        _, r2, d2, f2 = e2.step_host(a, obs_out=obs_h, auto_reset=True)
        assert np.array_equal(o1.cpu().numpy()[:, :obs_h.shape[1]], obs_h)
        assert np.array_equal(r1.cpu().numpy(), r2) and np.array_equal(d1.cpu().numpy(), d2)
        assert np.array_equal(f1.cpu().numpy(), f2)


@pytest.mark.parametrize('name,mode', [('d118_ac_random', 0), ('d118_ac_random', 1), ('d118_ac_random', 2),
                                       ('d118_ac_random', 3), ('d30_ac_random', 0), ('d30_ac_random', 1),
                                       ('d30_ac_random', 2), ('d14_ac_random', 1), ('d14_dc_random', 1),
                                       ('d14_dc_random', 2)])
def test_every_linear_solver_reproduces_the_reference(name, mode, monkeypatch):
    """The linear algebra behind B' / B'' / Bdc has four implementations (dense Gauss-Jordan inverse, sparse LDL^T +
    explicit inverses, sparse LDL^T + level-scheduled solves, hybrid sparse / dense top block; PpnStepArgs.sparse);
    the library picks one per grid size, PPN_SPARSE forces one.  Each must reproduce the reference fixture, including
    the steps whose node-splitting actions switch the factor to the two-rows-per-substation structure."""
    monkeypatch.setenv('PPN_SPARSE', str(mode))
    fx = Fixture(name)
    B = 2
    env = vec_env(fx, B)
    worst = 0.0
    for t in range(min(len(fx.actions), 40)):
        obs, reward, done, flag = env.step(np.repeat(fx.actions[t][None], B, axis=0))
        d, f = done.cpu().numpy(), flag.cpu().numpy()
        assert np.all(d == int(fx.done[t])) and np.all(f == int(fx.flag[t])), 'step %d: %s %s' % (t, d, f)
        if fx.done[t]:
            obs = env.process_game_over(done)
            expect = fx.reset_obs[t]
        else:
            expect = fx.obs[t]
        got = obs.cpu().numpy()
        worst = max(worst, float(np.max(np.abs(got[0] - expect))))
    assert worst < TOL, worst


def test_plan_switch_after_node_splitting_keeps_parity():
    """IEEE-118 handles start on the small hybrid shared-memory plan and move to the explicit-inverse plan once an env
    has applied a node switch (the kernel raises a flag in page-locked host memory): the trajectory must not notice."""
    fx = Fixture('d118_ac_random')
    B = 4
    env = vec_env(fx, B)
    c0 = env.counters()
    worst = 0.0
    for t in range(len(fx.actions)):
        obs, reward, done, flag = env.step(np.repeat(fx.actions[t][None], B, axis=0))
        assert np.all(done.cpu().numpy() == int(fx.done[t])) and np.all(flag.cpu().numpy() == int(fx.flag[t])), t
        if fx.done[t]:
            obs = env.process_game_over(done)
        expect = fx.reset_obs[t] if fx.done[t] else fx.obs[t]
        worst = max(worst, float(np.max(np.abs(obs.cpu().numpy()[B - 1] - expect))))
    assert worst < TOL, worst


@pytest.mark.parametrize('pinned', [True, False])
def test_chunked_host_entry_point_is_bit_identical(pinned):
    """ppn_step_host: page-locked result buffers are written by the kernel itself (zero-copy, one launch); pageable ones
    are staged with the batch cut into chunks on separate streams (copies overlap kernels).  Same bits either way as
    one device-pointer launch over the whole batch."""
    fx = Fixture('d14_ac_random')
    B = 2100
    rng = np.random.default_rng(5)
    nch = len(fx.chronics)
    start_c = rng.integers(0, nch, size=B).astype(np.int32)
    start_r = np.array([rng.integers(0, fx.chronics[c].n_rows - 1) for c in start_c], dtype=np.int32)
    e1 = vec_env(fx, B, start_chronics=start_c, start_rows=start_r)
    e2 = vec_env(fx, B, start_chronics=start_c, start_rows=start_r)
    case = fx.case
    nd = case.obs_dynamic_length
    act = torch.zeros((B, case.action_length), dtype=torch.uint8)
    if pinned:
        act = act.pin_memory()
    obs_h = np.zeros((B, nd + 3))                       # a row stride wider than the dynamic observation
    for t in range(12):
        a = np.zeros((B, case.action_length), dtype=np.uint8)
        rows = rng.integers(0, B, size=B // 4)
        a[rows, case.n_gen + case.n_load + 2 * case.n_line + rng.integers(case.n_line, size=len(rows))] = 1
        act.copy_(torch.from_numpy(a))
        o1, r1, d1, f1 = e1.step(a, auto_reset=True)
        if pinned:
            po, pr, pd, pf = e2.step_pinned(act, auto_reset=True)
            o2, r2, d2, f2 = po.numpy(), pr.numpy(), pd.numpy(), pf.numpy()
        else:
            _, r2, d2, f2 = e2.step_host(a, obs_out=obs_h, auto_reset=True)
            o2 = obs_h[:, :nd]
        assert np.array_equal(o1.cpu().numpy()[:, :nd], o2), t
        assert np.array_equal(r1.cpu().numpy(), r2) and np.array_equal(d1.cpu().numpy(), d2)
        assert np.array_equal(f1.cpu().numpy(), f2)
    if not pinned:
        assert e2.counters()['kernel_launches'] > e1.counters()['kernel_launches']   # one launch per chunk
    else:
        assert e2.counters()['kernel_launches'] == e1.counters()['kernel_launches']  # zero-copy: one launch


@pytest.mark.parametrize('drain', ['0', '1'])
@pytest.mark.parametrize('dtype', ['float64', 'float32'])
def test_pinned_rows_with_and_without_the_drain_kernel(monkeypatch, drain, dtype):
    """ppn_step_host with page-locked buffers delivers the observation rows either by stores of the step kernel itself or
    through device memory + the drain kernel (small rows; PPN_HOST_DRAIN forces either).  Same bits both ways; the row of an
    env that ended without auto-reset is not written at all."""
    monkeypatch.setenv('PPN_HOST_DRAIN', drain)
    fx = Fixture('d14_ac_random')
    B = 700
    rng = np.random.default_rng(11)
    start_r = rng.integers(0, 100, size=B).astype(np.int32)
    e1 = vec_env(fx, B, start_chronics=np.zeros(B, dtype=np.int32), start_rows=start_r)
    e2 = vec_env(fx, B, start_chronics=np.zeros(B, dtype=np.int32), start_rows=start_r)
    case = fx.case
    nd = case.obs_dynamic_length
    tdt = torch.float64 if dtype == 'float64' else torch.float32
    act = torch.zeros((B, case.action_length), dtype=torch.uint8).pin_memory()
    n_done = 0
    ended = np.zeros(B, dtype=bool)
    for t in range(10):
        a = np.zeros((B, case.action_length), dtype=np.uint8)
        rows = np.nonzero(rng.random(B) < 0.6)[0]
        a[rows, case.n_gen + case.n_load + 2 * case.n_line + rng.integers(case.n_line, size=len(rows))] = 1
        a[ended] = 0
        act.copy_(torch.from_numpy(a))
        o1, r1, d1, f1 = e1.step(a, auto_reset=False)
        if hasattr(e2, '_pin64' if dtype == 'float64' else '_pin32'):
            getattr(e2, '_pin64' if dtype == 'float64' else '_pin32')[0].fill_(-7.0)
        po, pr, pd, pf = e2.step_pinned(act, auto_reset=False, obs_dtype=tdt)
        d = d1.cpu().numpy().astype(bool)
        assert np.array_equal(d, pd.numpy().astype(bool)) and np.array_equal(f1.cpu().numpy(), pf.numpy())
        assert np.array_equal(r1.cpu().numpy(), pr.numpy())
        ref = o1.cpu().numpy()[:, :nd]
        if dtype == 'float32':
            ref = ref.astype(np.float32)
        live = ~d
        assert np.array_equal(ref[live], po.numpy()[live]), t
        if t > 0:
            assert np.all(po.numpy()[d] == -7.0), t
        n_done += int(d.sum())
        ended |= d
    assert n_done > 0


def test_drain_kernel_rounds_on_a_large_batch(monkeypatch):
    """More than 64 x 128 rows: the drain warps (at most 64) take their rows in rounds.  Same bits as the device entry point."""
    monkeypatch.setenv('PPN_HOST_DRAIN', '1')
    fx = Fixture('d14_ac_random')
    B = 9000
    rng = np.random.default_rng(4)
    start_r = rng.integers(0, 100, size=B).astype(np.int32)
    e1 = vec_env(fx, B, start_chronics=np.zeros(B, dtype=np.int32), start_rows=start_r)
    e2 = vec_env(fx, B, start_chronics=np.zeros(B, dtype=np.int32), start_rows=start_r)
    nd = fx.case.obs_dynamic_length
    act = torch.zeros((B, fx.case.action_length), dtype=torch.uint8).pin_memory()
    for t in range(4):
        o1, r1, d1, f1 = e1.step(act.numpy(), auto_reset=True)
        po, pr, pd, pf = e2.step_pinned(act, auto_reset=True)
        assert np.array_equal(o1.cpu().numpy()[:, :nd], po.numpy()), t
        assert np.array_equal(r1.cpu().numpy(), pr.numpy()) and np.array_equal(d1.cpu().numpy(), pd.numpy())
        assert np.array_equal(f1.cpu().numpy(), pf.numpy())


def test_float32_observation_rows_are_the_float64_ones_rounded_once():
    """ppn_step_host_f32 (VecRunEnv.step_pinned(obs_dtype=float32)): same trajectory, every observation value equal to the
    float64 one narrowed to float32 -- nothing else changes (rewards, done, flags stay float64 / integer)."""
    fx = Fixture('d14_ac_random')
    B = 40
    rng = np.random.default_rng(3)
    start_r = rng.integers(0, 100, size=B).astype(np.int32)
    e1 = vec_env(fx, B, start_chronics=np.zeros(B, dtype=np.int32), start_rows=start_r)
    e2 = vec_env(fx, B, start_chronics=np.zeros(B, dtype=np.int32), start_rows=start_r)
    nd = fx.case.obs_dynamic_length
    act = torch.zeros((B, fx.case.action_length), dtype=torch.uint8).pin_memory()
    for t in range(25):
        act.copy_(torch.from_numpy(np.repeat(fx.actions[t][None], B, axis=0)))
        o1, r1, d1, f1 = e1.step_pinned(act, auto_reset=True)
        o2, r2, d2, f2 = e2.step_pinned(act, auto_reset=True, obs_dtype=torch.float32)
        assert o2.dtype == torch.float32 and tuple(o2.shape) == (B, nd)
        assert np.array_equal(o1.numpy().astype(np.float32), o2.numpy()), t
        assert torch.equal(r1, r2) and torch.equal(d1, d2) and torch.equal(f1, f2)


def test_kernel_written_result_pack_equals_the_three_outputs():
    """ppn_set_result_pack: the row an env-sharded run all-gathers (reward[5] | done | flag) is written by the step
    kernel itself and must equal what packing the three output tensors gives."""
    from pypownet_b200 import sharding
    fx = Fixture('d14_ac_random')
    B = 6
    env = vec_env(fx, B)
    pack = env.enable_result_pack()
    for t in range(25):
        obs, reward, done, flag = env.step(np.repeat(fx.actions[t][None], B, axis=0), auto_reset=True)
        assert torch.equal(pack, sharding.pack_results(reward, done, flag)), t
        r, d, f = sharding.unpack_results(pack)
        assert torch.equal(d, done) and torch.equal(f, flag)


def test_is_action_valid_and_illegal_masks():
    fx = Fixture('d14_ac_random')
    env = vec_env(fx, 2)
    case = fx.case
    a = np.zeros((2, case.action_length), dtype=np.uint8)
    a[1, :] = 1                                         # far too many switches
    v = env.is_action_valid(a).cpu().numpy()
    assert v[0] and not v[1]
    env.step(a)
    ill = env.illegal.cpu().numpy()
    assert ill[1, 0] == 1 and ill[0].sum() == 0
    assert int(env.flag[1].item()) in (1, 2, 3, 4)
