"""Reward plug-in base class (pypownet/reward_signal.py:4-24) and the default five-term reward the shipped
environments define (parameters/default14/reward_signal.py:7-169), restated on arrays so that the same formula
runs on the device (csrc) and on the host."""
import numpy as np


class RewardSignal(object):
    """Template: compute_reward(observation, action, flag) -> list of sub-rewards."""

    def __init__(self):
        pass

    def compute_reward(self, observation, action, flag):
        return [0.]


class DefaultRewardConstants(object):
    """Hyper-parameters of the shipped CustomRewardSignal; `constant` is 14 / 30 / 118 in the shipped folders."""

    def __init__(self, constant):
        c = float(constant)
        self.constant = c
        self.line_usage = -1.
        self.distance_initial_grid = -.02
        self.loads_cut = -c / 5.
        self.prods_cut = -c / 10.
        self.loadflow_exception = -c
        self.illegal_switch = -c / 100.
        self.too_many_prods_cut = -c
        self.too_many_loads_cut = -c
        self.too_much_activated = -5 * c
        self.cost_line_switch = -.2
        self.cost_node_switch = -.1

    def as_array(self):
        return np.array([self.line_usage, self.distance_initial_grid, self.loads_cut, self.prods_cut,
                         self.loadflow_exception, self.illegal_switch, self.too_many_prods_cut,
                         self.too_many_loads_cut, self.too_much_activated, self.cost_line_switch,
                         self.cost_node_switch], dtype=np.float64)
