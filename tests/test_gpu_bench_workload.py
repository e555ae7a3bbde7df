"""CUDA vs the CPU oracle ON THE BENCHMARKED WORKLOADS themselves: `bench.build_workload` chronics and
limits, `bench.env_starts` offsets (env 0 = chronic 0, row 0), the do-nothing agent and the random node-splitting +
line-switching agent of `bench.random_action_bank`, restart on game over inside the step -- for IEEE-14 (BASELINE
configs[1]), IEEE-30 with the synthetic thermal limits that make the cascade fire (configs[2]) and IEEE-118
(configs[3], [4]).  Observation <= 1e-7 (north star: 1e-6 on voltages), done / flag / line status bit-exact."""
import numpy as np
import pytest

import bench
from oracle.flat import FlatEnv, Config

pytestmark = pytest.mark.gpu
TOL = 1e-7


@pytest.mark.parametrize('grid,agent,cascade,n_envs,steps', [
    ('case14', 'nothing', False, 64, 300), ('case14', 'random', False, 64, 150),
    ('case30', 'nothing', True, 32, 100), ('case30', 'random', True, 32, 100),
    ('case118', 'nothing', False, 32, 100), ('case118', 'random', False, 32, 100),
    ('case118', 'nothing', True, 32, 60),
])
def test_bench_workload_matches_the_oracle(grid, agent, cascade, n_envs, steps):
    from pypownet_b200.vec_env import VecRunEnv
    case, cfg, chronics, imaps = bench.build_workload(grid, cascade=cascade)
    B = n_envs
    # envs of the benchmarked batch (bench.shard_starts, 4096 per GPU): the first ones (env 0 = chronic 0, row 0) plus
    # a stretch far into the batch, and a few of rank 3 of an 8-GPU run
    sc0, sr0 = bench.shard_starts(4096, 0, 1)
    sc3, sr3 = bench.shard_starts(4096, 3, 8)
    pick = np.r_[np.arange(B // 2), 2900 + 7 * np.arange(B - B // 2 - 4)]
    sc = np.r_[sc0[pick], sc3[[5, 1700, 2811, 4001]]].astype(np.int32)
    sr = np.r_[sr0[pick], sr3[[5, 1700, 2811, 4001]]].astype(np.int32)
    assert sc[0] == 0 and sr[0] == 0 and len(sc) == B
    env = VecRunEnv(case, cfg, chronics, B, device=0, reward_constant=float(case.n_sub), thermal_limits=imaps,
                    start_chronics=sc, start_rows=sr)
    ocfg = Config(cfg, reward_constant=float(case.n_sub), n_sub=case.n_sub)
    refs = [FlatEnv(case, ocfg, chronics, start_id=int(sc[e]), thermal_limits=imaps, start_row=int(sr[e]))
            for e in range(B)]
    nd = case.obs_dynamic_length
    G, L, N = case.n_gen, case.n_load, case.n_line
    st0 = 4 * L + 4 * G + 3 * N                      # lines_status inside the observation vector
    got0 = env.obs.cpu().numpy()
    for e in range(B):
        assert np.max(np.abs(got0[e, :nd] - refs[e].observation_dynamic())) < TOL, e
    bank = bench.random_action_bank(case, B, seed=99) if agent == 'random' else None
    zeros = np.zeros((B, case.action_length), dtype=np.uint8)
    worst, n_done, depth_seen = 0.0, 0, 0
    for t in range(steps):
        acts = bank[t % len(bank)] if bank is not None else zeros
        obs, reward, done, flag = env.step(acts, auto_reset=True)
        got, r, d, f = obs.cpu().numpy(), reward.cpu().numpy(), done.cpu().numpy(), flag.cpu().numpy()
        for e in range(B):
            o2, r2, d2, f2, _ = refs[e].step(acts[e])
            depth_seen = max(depth_seen, refs[e].last_depth)
            assert (bool(d[e]), int(f[e])) == (bool(d2), int(f2)), 'step %d env %d' % (t, e)
            assert np.max(np.abs(r[e] - r2)) < TOL, 'step %d env %d' % (t, e)
            if d2:
                o2 = refs[e].process_game_over()
                n_done += 1
            assert np.array_equal(got[e, st0:st0 + N], o2[st0:st0 + N]), 'step %d env %d: line status' % (t, e)
            err = float(np.max(np.abs(got[e, :nd] - o2)))
            assert err < TOL, 'step %d env %d: %g' % (t, e, err)
            worst = max(worst, err)
    assert n_done > 0
    if cascade:
        assert depth_seen >= 2          # lines did trip and the load-flow was re-run within a step
    print('%s %s cascade=%s: %d envs x %d steps, %d game overs, max cascade depth %d, max |cuda - oracle| = %.3g'
          % (grid, agent, cascade, B, steps, n_done, depth_seen, worst))
