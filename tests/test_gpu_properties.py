"""Size-independent properties of the CUDA step path at BASELINE.json batch sizes (no oracle in the loop):
determinism, batch-size independence, `simulate` leaves no trace, candidate batching, state get/set round trip,
masked process_game_over, chronic looping modes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def workload(grid='case14'):
    import bench
    return bench.build_workload(grid)


def make(grid, B, **kw):
    import bench
    from pypownet_b200.vec_env import VecRunEnv
    case, cfg, chronics, imaps = workload(grid)
    sc, sr = bench.env_starts(B)
    return VecRunEnv(case, cfg, chronics, B, reward_constant=float(case.n_sub), thermal_limits=imaps,
                     start_chronics=sc, start_rows=sr, **kw), case


@pytest.mark.parametrize('grid,B', [('case14', 4096), ('case30', 2048), ('case118', 296)])
def test_deterministic_and_batch_size_independent(grid, B):
    from pypownet_b200.agent import VecRandomSplitAndSwitch
    e1, case = make(grid, B)
    e2, _ = make(grid, B)
    small, _ = make(grid, 64)
    agent = VecRandomSplitAndSwitch(e1, seed=5)
    for t in range(12):
        a = agent.act()
        o1, r1, d1, f1 = [x.clone() for x in e1.step(a, auto_reset=True)]
        o2, r2, d2, f2 = e2.step(a, auto_reset=True)
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2) and torch.equal(f1, f2)
        o3, r3, d3, f3 = small.step(a[:64].contiguous(), auto_reset=True)
        assert torch.equal(o1[:64], o3) and torch.equal(r1[:64], r3) and torch.equal(f1[:64], f3)
    assert bool(torch.isfinite(o1).all())
    c = e1.counters()
    assert c['env_steps'] == 13 * B and c['loadflows'] >= c['env_steps']      # 12 steps + the initial cascade


def test_simulate_leaves_no_trace_and_batches_candidates():
    from pypownet_b200 import _lib
    from pypownet_b200.agent import VecRandomSplitAndSwitch
    env, case = make('case14', 512)
    agent = VecRandomSplitAndSwitch(env, seed=9)
    for t in range(5):
        env.step(agent.act(), auto_reset=True)
    before = [env.get_state(f).clone() for f in (_lib.STATE_REAL, _lib.STATE_TOPOLOGY, _lib.STATE_COUNTERS)]
    K = 3
    cand = torch.stack([agent.act() for _ in range(K)], dim=1).reshape(512 * K, case.action_length).contiguous()
    so, sr, sd, sf = env.simulate(cand, n_candidates=K)
    after = [env.get_state(f) for f in (_lib.STATE_REAL, _lib.STATE_TOPOLOGY, _lib.STATE_COUNTERS)]
    for b, a in zip(before, after):
        assert torch.equal(a, b)                                   # test_simulate.py:339-536: no trace
    # candidate k of env e == a single-candidate simulate of that action
    for k in range(K):
        o1, r1, d1, f1 = env.simulate(cand.reshape(512, K, -1)[:, k].contiguous())
        assert torch.equal(sd.reshape(512, K)[:, k], d1) and torch.equal(sf.reshape(512, K)[:, k], f1)
        live = ~d1.bool()
        assert torch.equal(so.reshape(512, K, -1)[:, k][live], o1[live])
        assert torch.equal(sr.reshape(512, K, 5)[:, k], r1)
    # a do-nothing simulate reports the planned injections of the current row as its loads
    so, sr, sd, sf = env.simulate(torch.zeros((512, case.action_length), dtype=torch.uint8))
    L = case.n_load
    live = ~sd.bool()
    assert torch.equal(so[live][:, :L], env.obs[live][:, 2 * L:3 * L])         # test_simulate.py:273-326, exactly 0.0


def test_state_round_trip_and_masked_game_over():
    from pypownet_b200 import _lib
    env, case = make('case14', 128)
    for t in range(6):
        env.step(None, auto_reset=True)
    snap = [env.get_state(f).clone() for f in (_lib.STATE_REAL, _lib.STATE_TOPOLOGY, _lib.STATE_COUNTERS)]
    o1 = env.step(None, auto_reset=False)[0].clone()
    d1, f1 = env.done.clone(), env.flag.clone()
    for f, s in zip((_lib.STATE_REAL, _lib.STATE_TOPOLOGY, _lib.STATE_COUNTERS), snap):
        env.set_state(f, s)
    o2 = env.step(None, auto_reset=False)[0]
    assert torch.equal(env.done, d1) and torch.equal(env.flag, f1)
    live = ~d1.bool()
    assert torch.equal(o1[live], o2[live])
    # process_game_over only touches the masked envs
    topo_before = env.get_state(_lib.STATE_TOPOLOGY).clone()
    mask = d1.clone()
    env.process_game_over(mask)
    topo_after = env.get_state(_lib.STATE_TOPOLOGY)
    assert torch.equal(topo_after[~mask.bool()], topo_before[~mask.bool()])
    N, G, L = case.n_line, case.n_gen, case.n_load
    if mask.any():
        assert int(topo_after[mask.bool()][:, :G + L + 2 * N].sum().item()) == 0     # initial topology restored


@pytest.mark.parametrize('mode', ['natural', 'fixed', 'random'])
def test_chronic_looping_modes(mode):
    import bench
    from pypownet_b200 import _lib
    from pypownet_b200.vec_env import VecRunEnv
    case, cfg, chronics, imaps = workload('case14')
    short = chronics[:4]
    B = 32
    start_c = np.arange(B, dtype=np.int32) % 4
    start_r = np.full(B, short[0].n_rows - 3, dtype=np.int32)
    envs = [VecRunEnv(case, cfg, short, B, reward_constant=14., thermal_limits=imaps, loop_mode=mode, seed=7,
                      start_chronics=start_c, start_rows=start_r) for _ in range(2)]
    for t in range(6):
        for e in envs:
            e.step(None, auto_reset=True)
    cur = [e.get_state(_lib.STATE_COUNTERS)[:, -4:].cpu().numpy() for e in envs]
    assert np.array_equal(cur[0], cur[1])                          # reproducible, also for the hashed 'random' mode
    chronic_now = cur[0][:, 0]
    if mode == 'natural':
        assert np.array_equal(chronic_now, (start_c + 1) % 4)
    elif mode == 'fixed':
        assert np.array_equal(chronic_now, start_c)
    else:
        assert chronic_now.min() >= 0 and chronic_now.max() < 4
    assert (cur[0][:, 1] < 8).all()                                # restarted near the beginning of a chronic
