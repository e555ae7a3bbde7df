"""Runner: the act -> step -> (process_game_over) -> feed_reward loop (reference: pypownet/runner.py:26-145), with
the same `runner.log` / `machine_logs.csv` outputs, plus VecRunner for batches."""
import csv
import datetime
import logging
import os


class Runner(object):
    def __init__(self, environment, agent, render=False, verbose=False, vverbose=False, parameters=None, level=None,
                 max_iter=None, log_filepath='runner.log', machinelog_filepath='machine_logs.csv'):
        self.environment, self.agent = environment, agent
        self.verbose, self.render = verbose, render
        self.parameters, self.level, self.max_iter = parameters, level, max_iter
        self.logger = logging.getLogger('pypownet')
        self.logger.setLevel(logging.DEBUG)
        if log_filepath is not None and not any(isinstance(h, logging.FileHandler) for h in self.logger.handlers):
            fh = logging.FileHandler(filename=log_filepath, mode='w+')
            fh.setLevel(logging.DEBUG)
            fh.setFormatter(logging.Formatter('%(asctime)s - %(name)s - %(levelname)s - %(message)s'))
            self.logger.addHandler(fh)
        if verbose or vverbose:
            ch = logging.StreamHandler()
            ch.setLevel(logging.DEBUG if vverbose else logging.INFO)
            self.logger.addHandler(ch)
        self.csv_writer = None
        if machinelog_filepath is not None:
            self._csv_file = open(machinelog_filepath, 'w')
            self.csv_writer = csv.writer(self._csv_file, delimiter=';')
            self.csv_writer.writerow(['param_env_name', 'level', 'chronic_name', 'max_iter', 'timestep', 'time',
                                      'game_over', 'timestep_reward_aslist', 'timestep_reward', 'cumulated_reward'])

    def step(self, observation):
        """One RL step (runner.py:72-103): returns (observation, action, reward, reward_aslist, done)."""
        action = self.agent.act(observation)
        observation, reward_aslist, done, info = self.environment.step(action, do_sum=False)
        if done:
            self.logger.warning('GAME OVER! Resetting grid... (hint: %s)' % info.text)
            observation = self.environment.process_game_over()
        elif info:
            self.logger.warning(info.text)
        reward = sum(reward_aslist)
        self.agent.feed_reward(action, observation, reward_aslist)
        return observation, action, reward, reward_aslist, done

    def loop(self, iterations, epochs=1):
        cumul_rew = 0.0
        for _ in range(epochs):
            self.logger.warning('Resetting environment...')
            observation = self.environment.reset()
            for i_iter in range(1, iterations + 1):
                observation, action, reward, reward_aslist, done = self.step(observation)
                cumul_rew += reward
                self.logger.info('step %d/%d - reward: %.2f; cumulative reward: %.2f' % (i_iter, iterations, reward,
                                                                                         cumul_rew))
                self.dump_machinelogs(i_iter, done, reward, reward_aslist, cumul_rew,
                                      self.environment.get_current_datetime())
        return cumul_rew

    def dump_machinelogs(self, timestep_id, done, reward, reward_aslist, cumul_rew, datetime_):
        if self.csv_writer is None:
            return
        self.csv_writer.writerow([self.parameters, self.level, self.environment.get_current_chronic_name(),
                                  self.max_iter, timestep_id, datetime_.strftime('%Y-%m-%d %H:%M'), done,
                                  reward_aslist, reward, cumul_rew])
        self._csv_file.flush()


class VecRunner(object):
    """Batched loop: every env of a VecRunEnv acts, steps and restarts on game over inside one kernel launch."""

    def __init__(self, vec_env, vec_agent):
        self.env, self.agent = vec_env, vec_agent

    def loop(self, iterations):
        """Returns (cumulative reward per env [B] tensor, number of game overs per env [B] tensor)."""
        import torch
        env = self.env
        cum = torch.zeros((env.n_envs,), dtype=torch.float64, device=env.device)
        overs = torch.zeros((env.n_envs,), dtype=torch.int64, device=env.device)
        obs = env.obs
        for _ in range(iterations):
            obs, reward, done, flag = env.step(self.agent.act(obs), auto_reset=True)
            cum += reward.sum(dim=1)
            overs += done.to(torch.int64)
        return cum, overs
