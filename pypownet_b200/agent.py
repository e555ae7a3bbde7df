"""Agents on either side of the step path (reference: pypownet/agent.py:9-158).

Same contract as the reference: `Agent(environment).act(observation) -> action`, `feed_reward(action,
consequent_observation, rewards_aslist)`.  Single-env agents work on a RunEnv; the Vec* agents produce one action row
per env of a VecRunEnv directly on the GPU (what BASELINE.json's batched configurations need)."""
import numpy as np


class Agent(object):
    def __init__(self, environment):
        self.environment = environment

    def act(self, observation):
        """observation: array (RunEnv convention) or Observation -> an action (array or Action)."""
        return self.environment.action_space.get_do_nothing_action()

    def feed_reward(self, action, consequent_observation, rewards_aslist):
        pass


class DoNothing(Agent):
    def act(self, observation):
        return self.environment.action_space.get_do_nothing_action(as_class_Action=True)


class RandomAction(Agent):
    """Uniformly random switch vector (agent.py:41-56); nearly always rejected as too many activations."""

    def act(self, observation):
        return self.environment.action_space.sample()


class RandomLineSwitch(Agent):
    """Switches the status of one random line per timestep (agent.py:78-111)."""

    def act(self, observation):
        space = self.environment.action_space
        action = space.get_do_nothing_action(as_class_Action=True)
        space.set_lines_status_switch_from_id(action=action, line_id=np.random.randint(space.lines_status_subaction_length),
                                              new_switch_value=1)
        return action


class RandomNodeSplitting(Agent):
    """One random substation, random new configuration of its elements (agent.py:116-158)."""

    def act(self, observation):
        space = self.environment.action_space
        action = space.get_do_nothing_action(as_class_Action=True)
        sub = np.random.choice(space.substations_ids)
        size = space.get_number_elements_of_substation(sub)
        target = np.random.choice([0, 1], size=(size,))
        space.set_substation_switches_in_action(action=action, substation_id=sub, new_values=target)
        current, _ = space.get_substation_switches_in_action(action, sub)
        assert np.all(current == target)
        return action


class VecDoNothing(object):
    """Batched do-nothing agent: a zero action tensor that lives on the env's GPU."""

    def __init__(self, vec_env):
        import torch
        self.actions = torch.zeros((vec_env.n_envs, vec_env.action_length), dtype=torch.uint8, device=vec_env.device)

    def act(self, observations=None):
        return self.actions


class VecRandomSplitAndSwitch(object):
    """RandomNodeSplitting U RandomLineSwitch per env per step: one substation chosen
    uniformly with its element switches i.i.d. Bernoulli(1/2), plus one uniformly chosen line-status switch.
    Generated on the device from a torch.Generator, so the stream is reproducible for a given seed and device."""

    def __init__(self, vec_env, seed=0):
        import torch
        self.env = vec_env
        case = vec_env.case
        self.gen = torch.Generator(device=vec_env.device)
        self.gen.manual_seed(seed)
        self.elem_sub = torch.from_numpy(case.elem_sub.astype(np.int64)).to(vec_env.device)
        self.n_node = case.n_gen + case.n_load + 2 * case.n_line

    def act(self, observations=None):
        import torch
        env, case = self.env, self.env.case
        B, dev = env.n_envs, env.device
        sub = torch.randint(0, case.n_sub, (B, 1), generator=self.gen, device=dev)
        bits = torch.randint(0, 2, (B, self.n_node), generator=self.gen, device=dev, dtype=torch.uint8)
        nodes = bits * (self.elem_sub[None, :] == sub).to(torch.uint8)
        line = torch.randint(0, case.n_line, (B,), generator=self.gen, device=dev)
        lines = torch.zeros((B, case.n_line), dtype=torch.uint8, device=dev)
        lines[torch.arange(B, device=dev), line] = 1
        return torch.cat([nodes, lines], dim=1)


class VecGreedySearch(object):
    """Batched GreedySearch (reference: pypownet/agent.py:227-325): a depth-1 tree search that simulates the do-nothing
    action, every single line-status switch and, for every substation with 4 or 5 elements, every switch pattern of its
    elements whose first switch is 0, and plays the candidate with the highest summed reward (the first one on ties,
    as `rewards.index(max(rewards))`).  The reference runs its 1 + N + sum 2^(k-1) simulations one after the other
    (69 for IEEE-14); here all candidates of all envs are ONE ppn_simulate launch (B x K rows, no state change)."""

    def __init__(self, vec_env):
        import itertools
        import torch
        self.env = vec_env
        case = vec_env.case
        A, n_node = case.action_length, case.n_gen + case.n_load + 2 * case.n_line
        cands = [np.zeros(A, dtype=np.uint8)]
        for l in range(case.n_line):
            a = np.zeros(A, dtype=np.uint8)
            a[n_node + l] = 1
            cands.append(a)
        for s in range(case.n_sub):
            el = np.flatnonzero(case.elem_sub == s)          # productions, loads, line origins, line extremities
            if 3 < len(el) < 6:
                for conf in itertools.product([0, 1], repeat=len(el) - 1):
                    a = np.zeros(A, dtype=np.uint8)
                    a[el] = (0,) + conf
                    cands.append(a)
        self.candidates = torch.from_numpy(np.array(cands)).to(vec_env.device)          # [K, A]
        self.n_candidates = len(cands)
        self.last_rewards = None

    def act(self, observations=None):
        import torch
        env, K = self.env, self.n_candidates
        batch = self.candidates.unsqueeze(0).expand(env.n_envs, K, -1).reshape(env.n_envs * K, -1).contiguous()
        _, reward, done, flag = env.simulate(batch, n_candidates=K)
        r = reward.view(env.n_envs, K, 5)
        total = (((r[..., 0] + r[..., 1]) + r[..., 2]) + r[..., 3]) + r[..., 4]       # sum(reward_aslist), same order
        self.last_rewards, self.last_done, self.last_flag = r, done.view(env.n_envs, K), flag.view(env.n_envs, K)
        best = torch.argmax((total == total.max(dim=1, keepdim=True).values).to(torch.uint8), dim=1)   # first maximum
        return self.candidates[best]
