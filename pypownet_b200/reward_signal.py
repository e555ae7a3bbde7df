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


class DefaultRewardSignal(RewardSignal):
    """The five-term reward of the shipped environments (parameters/default14/reward_signal.py:7-169), host version:
    [load cut, production cut, action cost, distance to the reference grid, line usage].  The step kernel computes
    the same five numbers on the device; this class is what a reward_signal.py plug-in can subclass."""

    def __init__(self, constant=14):
        super(DefaultRewardSignal, self).__init__()
        k = DefaultRewardConstants(constant)
        self.__dict__.update({
            'multiplicative_factor_line_usage_reward': k.line_usage,
            'multiplicative_factor_distance_initial_grid': k.distance_initial_grid,
            'multiplicative_factor_number_loads_cut': k.loads_cut,
            'multiplicative_factor_number_prods_cut': k.prods_cut,
            'connexity_exception_reward': k.loadflow_exception, 'loadflow_exception_reward': k.loadflow_exception,
            'multiplicative_factor_number_illegal_lines_reconnection': k.illegal_switch,
            'too_many_productions_cut': k.too_many_prods_cut, 'too_many_consumptions_cut': k.too_many_loads_cut,
            'multiplicative_factor_number_line_switches': k.cost_line_switch,
            'multiplicative_factor_number_node_switches': k.cost_node_switch})
        self._k = k

    def compute_reward(self, observation, action, flag):
        from pypownet_b200 import environment as E
        k = self._k
        cost = k.cost_node_switch * float(np.sum(action.get_node_splitting_subaction())) + \
            k.cost_line_switch * float(np.sum(action.get_lines_status_subaction()))
        if isinstance(flag, E.DivergingLoadflowException):
            return [0., 0., cost, k.loadflow_exception, 0.]
        if isinstance(flag, E.TooManyProductionsCut):
            return [0., k.too_many_prods_cut, 0., 0., 0.]
        if isinstance(flag, E.TooManyConsumptionsCut):
            return [k.too_many_loads_cut, 0., 0., 0., 0.]
        o = observation
        dist = float(np.sum(o.productions_nodes) + np.sum(o.loads_nodes) + np.sum(o.lines_or_nodes) +
                     np.sum(o.lines_ex_nodes))
        usage = np.asarray(o.ampere_flows) / np.asarray(o.thermal_limits)
        r = [k.loads_cut * float(np.sum(o.are_loads_cut)), k.prods_cut * float(np.sum(o.are_productions_cut)), cost,
             k.distance_initial_grid * dist, k.line_usage * float(np.sum(np.square(usage)))]
        if isinstance(flag, E.IllegalActionException):
            if flag.has_too_much_activations:
                r[2] += k.too_much_activated
            else:
                r[2] += k.illegal_switch * float(np.sum(flag.illegal_lines_reconnections) +
                                                 np.sum(flag.illegal_unavailable_lines_switches) +
                                                 np.sum(flag.illegal_oncoolown_substations_switches))
        return r
