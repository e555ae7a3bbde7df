"""The RunEnv drop-in (pypownet_b200.environment) against fixtures recorded through the reference's RunEnv: same
tuple (observation array | None, reward list, done, flag INSTANCE of the reference's exception classes)."""
import numpy as np
import pytest

from golden_util import Fixture, write_environment_folder

pytestmark = pytest.mark.gpu
TOL = 1e-7


@pytest.mark.parametrize('name', ['d14_tests_basic', 'd14_tests_hard_overflow', 'd14_ac_random'])
def test_runenv_reproduces_reference_run(name, tmp_path):
    from pypownet_b200 import environment as E
    fx = Fixture(name)
    folder = write_environment_folder(fx, str(tmp_path / 'env'))
    env = E.RunEnv(folder, 'level0', game_over_mode=fx.mode)
    assert env.action_space.action_length == fx.case.action_length
    assert np.max(np.abs(env.get_observation() - fx.obs0)) < TOL
    classes = {0: type(None), 1: E.IllegalActionException, 2: E.DivergingLoadflowException,
               3: E.TooManyConsumptionsCut, 4: E.TooManyProductionsCut}
    for t in range(min(len(fx.actions), 60)):
        if fx.has_sim:
            so, sr, sd, sf = env.simulate(fx.sim_actions[t], do_sum=False)
            assert sd == bool(fx.sim_done[t]) and isinstance(sf, classes[int(fx.sim_flag[t])])
            if not sd:
                assert np.max(np.abs(so - fx.sim_obs[t])) < TOL
        obs, reward, done, flag = env.step(fx.actions[t], do_sum=False)
        assert done == bool(fx.done[t]) and isinstance(flag, classes[int(fx.flag[t])]), t
        if fx.default_reward:
            assert np.max(np.abs(np.asarray(reward) - fx.reward[t])) < TOL
        if done:
            assert obs is None
            obs = env.process_game_over()
            assert np.max(np.abs(obs - fx.reset_obs[t])) < TOL
        else:
            assert np.max(np.abs(obs - fx.obs[t])) < TOL
            o = env.observation_space.array_to_observation(obs)
            assert np.array_equal(o.as_array(), obs)                      # ObsToArrayAndBack, test_core.py:68-75
    with pytest.raises(ValueError):
        env.step(None)
    with pytest.raises(ValueError):
        env.step(np.zeros(3))


def test_runner_and_agents(tmp_path):
    from pypownet_b200 import environment as E
    from pypownet_b200.agent import DoNothing, RandomLineSwitch, RandomNodeSplitting
    from pypownet_b200.runner import Runner
    fx = Fixture('d14_ac_nothing')
    folder = write_environment_folder(fx, str(tmp_path / 'env'))
    env = E.RunEnv(folder, 'level0')
    for agent_class in (DoNothing, RandomLineSwitch, RandomNodeSplitting):
        runner = Runner(env, agent_class(env), log_filepath=None, machinelog_filepath=str(tmp_path / 'm.csv'))
        total = runner.loop(iterations=15)
        assert np.isfinite(total)
    # the do-nothing run follows the recorded reference run
    env = E.RunEnv(folder, 'level0')
    runner = Runner(env, DoNothing(env), log_filepath=None, machinelog_filepath=None)
    obs = env.get_observation()
    for t in range(25):
        obs, action, reward, reward_aslist, done = runner.step(obs)
        assert done == bool(fx.done[t])
        expect = fx.reset_obs[t] if done else fx.obs[t]
        assert np.max(np.abs(obs - expect)) < TOL
        assert abs(reward - fx.reward[t].sum()) < TOL


def test_vec_runner_random_agent_runs():
    from pypownet_b200.agent import VecRandomSplitAndSwitch
    from pypownet_b200.runner import VecRunner
    from pypownet_b200.vec_env import VecRunEnv
    fx = Fixture('d118_ac_random')
    env = VecRunEnv(fx.case, fx.config, fx.chronics, 16, reward_constant=fx.reward_constant,
                    thermal_limits=fx.thermal_limits)
    cum, overs = VecRunner(env, VecRandomSplitAndSwitch(env, seed=3)).loop(20)
    assert np.all(np.isfinite(cum.cpu().numpy())) and int(overs.sum().item()) >= 0


def test_batched_greedy_search_reproduces_the_reference_agent():
    """tests/golden/greedy/d14_greedy.npz holds what the reference GreedySearch (agent.py:227-325) simulated on the
    reference environment: 69 candidate actions per step with their five sub-rewards, the action it chose and the
    outcome of playing it.  VecGreedySearch evaluates all candidates of all envs in one ppn_simulate launch."""
    import numpy as np
    import torch
    from golden_util import Fixture
    from pypownet_b200.agent import VecGreedySearch
    from pypownet_b200.vec_env import VecRunEnv
    fx = Fixture('greedy/d14_greedy')
    z = fx.z
    B = 3
    env = VecRunEnv(fx.case, fx.config, fx.chronics, B, device=0, game_over_mode=fx.mode,
                    reward_constant=fx.reward_constant, thermal_limits=fx.thermal_limits)
    agent = VecGreedySearch(env)
    assert agent.n_candidates == z['cand_actions'].shape[1]
    assert np.array_equal(agent.candidates.cpu().numpy(), z['cand_actions'][0])       # same candidates, same order
    for t in range(len(fx.actions)):
        a = agent.act()
        got = agent.last_rewards.cpu().numpy()
        for e in (0, B - 1):
            assert np.array_equal(agent.last_done[e].cpu().numpy().astype(bool), z['cand_done'][t])
            assert np.array_equal(agent.last_flag[e].cpu().numpy(), z['cand_flag'][t])
            assert np.nanmax(np.abs(got[e] - z['cand_reward'][t])) < 1e-7
        assert np.array_equal(a.cpu().numpy(), np.repeat(fx.actions[t][None], B, axis=0))
        obs, reward, done, flag = env.step(a)
        assert not bool(done.any()) and not fx.done[t]
        assert np.max(np.abs(obs.cpu().numpy()[0] - fx.obs[t])) < 1e-7
        assert np.max(np.abs(reward.cpu().numpy() - fx.reward[t][None])) < 1e-7
