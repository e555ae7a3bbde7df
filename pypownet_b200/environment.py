"""Host-side mirror of pypownet.environment for the step path (reference: pypownet/environment.py).

Same names, argument meaning and error behaviour as the reference for what sits either side of the hot path:
`RunEnv(parameters_folder, game_level, ...)` with `reset / step / simulate / process_game_over / get_observation /
is_action_valid`, `ActionSpace`, `ObservationSpace`, `Observation` (with `as_array`, whose layout is the device
packing contract) and the four in-step exception classes that `step` RETURNS as its flag (game.py:861-885).
The arithmetic is not here: RunEnv drives a one-env VecRunEnv, i.e. the CUDA library behind the C ABI.
"""
from collections import OrderedDict
from enum import Enum

import numpy as np

from pypownet_b200 import _lib


# ----------------------------------------------------------------------------------------------------- exceptions
class DivergingLoadflowException(Exception):
    def __init__(self, last_observation, *args):
        super(DivergingLoadflowException, self).__init__(last_observation, *args)
        self.last_observation = last_observation
        self.text = args[0] if args else None


class TooManyProductionsCut(Exception):
    def __init__(self, *args):
        super(TooManyProductionsCut, self).__init__(*args)
        self.text = args[0] if args else None


class TooManyConsumptionsCut(Exception):
    def __init__(self, *args):
        super(TooManyConsumptionsCut, self).__init__(*args)
        self.text = args[0] if args else None


class IllegalActionException(Exception):
    """environment.py:14-20 / game.py:24-56: carries the three illegality masks and the too-many-activations bit."""

    def __init__(self, text, has_too_much_activations, illegal_lines_reconnections=None,
                 illegal_unavailable_lines_switches=None, illegal_oncoolown_substations_switches=None, *args):
        super(IllegalActionException, self).__init__(text, *args)
        self.text = text
        self.has_too_much_activations = has_too_much_activations
        self.illegal_lines_reconnections = illegal_lines_reconnections
        self.illegal_unavailable_lines_switches = illegal_unavailable_lines_switches
        self.illegal_oncoolown_substations_switches = illegal_oncoolown_substations_switches

    def get_has_too_much_activations(self):
        return self.has_too_much_activations

    def get_illegal_broken_lines_reconnections(self):
        return self.illegal_lines_reconnections

    def get_illegal_oncoolown_lines_switches(self):
        return self.illegal_unavailable_lines_switches

    def get_illegal_oncoolown_substations_switches(self):
        return self.illegal_oncoolown_substations_switches

    @property
    def is_empty(self):
        if self.has_too_much_activations:
            return False
        return not (np.any(self.illegal_lines_reconnections) or np.any(self.illegal_unavailable_lines_switches) or
                    np.any(self.illegal_oncoolown_substations_switches))


class ElementType(Enum):
    PRODUCTION = "production"
    CONSUMPTION = "consumption"
    ORIGIN_POWER_LINE = "origin of power line"
    EXTREMITY_POWER_LINE = "extremity of power line"


# --------------------------------------------------------------------------------------------------------- action
class Action(object):
    """The five switch sub-vectors of an action (game.py:74-251)."""

    def __init__(self, prods_switches_subaction, loads_switches_subaction, lines_or_switches_subaction,
                 lines_ex_switches_subaction, lines_status_subaction):
        self.prods_switches_subaction = np.asarray(prods_switches_subaction).astype(int)
        self.loads_switches_subaction = np.asarray(loads_switches_subaction).astype(int)
        self.lines_or_switches_subaction = np.asarray(lines_or_switches_subaction).astype(int)
        self.lines_ex_switches_subaction = np.asarray(lines_ex_switches_subaction).astype(int)
        self.lines_status_subaction = np.asarray(lines_status_subaction).astype(int)

    def get_node_splitting_subaction(self):
        return np.concatenate((self.prods_switches_subaction, self.loads_switches_subaction,
                               self.lines_or_switches_subaction, self.lines_ex_switches_subaction))

    def get_lines_status_subaction(self):
        return self.lines_status_subaction

    def as_array(self):
        return np.concatenate((self.get_node_splitting_subaction(), self.lines_status_subaction))

    def __len__(self, do_sum=True):
        sizes = (len(self.prods_switches_subaction), len(self.loads_switches_subaction),
                 len(self.lines_or_switches_subaction), len(self.lines_ex_switches_subaction),
                 len(self.lines_status_subaction))
        return sum(sizes) if do_sum else sizes

    def __getitem__(self, item):
        return self.as_array()[item]


class ActionSpace(object):
    """environment.py:46-274 without the gym base class: a binary vector of length G + L + 3N laid out as
    prods | loads | lines origin | lines extremity node switches, then line status switches."""

    def __init__(self, number_generators, number_consumers, number_power_lines, number_substations, substations_ids,
                 prods_subs_ids, loads_subs_ids, lines_or_subs_id, lines_ex_subs_id):
        self.prods_switches_subaction_length = number_generators
        self.loads_switches_subaction_length = number_consumers
        self.lines_or_switches_subaction_length = number_power_lines
        self.lines_ex_switches_subaction_length = number_power_lines
        self.lines_status_subaction_length = number_power_lines
        self.action_length = number_generators + number_consumers + 3 * number_power_lines
        self.n = self.action_length
        self.shape = (self.action_length,)
        self.substations_ids = np.asarray(substations_ids)
        self.prods_subs_ids = np.asarray(prods_subs_ids)
        self.loads_subs_ids = np.asarray(loads_subs_ids)
        self.lines_or_subs_id = np.asarray(lines_or_subs_id)
        self.lines_ex_subs_id = np.asarray(lines_ex_subs_id)
        self._substations_n_elements = [
            int((self.prods_subs_ids == s).sum() + (self.loads_subs_ids == s).sum() +
                (self.lines_or_subs_id == s).sum() + (self.lines_ex_subs_id == s).sum()) for s in self.substations_ids]

    def sample(self):
        return np.random.randint(0, 2, size=self.action_length)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == (self.action_length,) and bool(np.all((x == 0) | (x == 1)))

    def get_do_nothing_action(self, as_class_Action=False):
        a = Action(np.zeros(self.prods_switches_subaction_length), np.zeros(self.loads_switches_subaction_length),
                   np.zeros(self.lines_or_switches_subaction_length), np.zeros(self.lines_ex_switches_subaction_length),
                   np.zeros(self.lines_status_subaction_length))
        return a if as_class_Action else a.as_array()

    def array_to_action(self, array):
        if isinstance(array, Action):
            return array
        if len(array) != self.action_length:
            raise ValueError('Expected action as a binary array of length %d, got %d' % (self.action_length,
                                                                                        len(array)))
        array = np.asarray(array)
        G, L, N = self.prods_switches_subaction_length, self.loads_switches_subaction_length, \
            self.lines_status_subaction_length
        return Action(array[:G], array[G:G + L], array[G + L:G + L + N], array[G + L + N:G + L + 2 * N],
                      array[G + L + 2 * N:])

    def _verify_action_shape(self, action):
        if action is None:
            raise ValueError('Expected binary array of length %d, got None' % self.action_length)
        if isinstance(action, Action):
            a = Action(action.prods_switches_subaction.copy(), action.loads_switches_subaction.copy(),
                       action.lines_or_switches_subaction.copy(), action.lines_ex_switches_subaction.copy(),
                       action.lines_status_subaction.copy())
        else:
            a = self.array_to_action(action)
        sizes = a.__len__(do_sum=False)
        expected = (self.prods_switches_subaction_length, self.loads_switches_subaction_length,
                    self.lines_or_switches_subaction_length, self.lines_ex_switches_subaction_length,
                    self.lines_status_subaction_length)
        names = ('prods_switches_subaction', 'loads_switches_subaction', 'lines_or_switches_subaction',
                 'lines_ex_subaction', 'lines_status_subaction')
        for got, exp, nm in zip(sizes, expected, names):
            if got and got != exp:
                raise ValueError('Expected %s subaction of size %d, got %d' % (nm, exp, got))
        return a

    def get_number_elements_of_substation(self, substation_id):
        assert substation_id in self.substations_ids
        return self._substations_n_elements[int(np.where(self.substations_ids == substation_id)[0][0])]

    def get_substation_switches_in_action(self, action, substation_id, concatenated_output=True):
        action = self.array_to_action(action)
        assert substation_id in self.substations_ids, 'Substation with id %d does not exist' % substation_id
        parts = (action.prods_switches_subaction[self.prods_subs_ids == substation_id],
                 action.loads_switches_subaction[self.loads_subs_ids == substation_id],
                 action.lines_or_switches_subaction[self.lines_or_subs_id == substation_id],
                 action.lines_ex_switches_subaction[self.lines_ex_subs_id == substation_id])
        kinds = (ElementType.PRODUCTION, ElementType.CONSUMPTION, ElementType.ORIGIN_POWER_LINE,
                 ElementType.EXTREMITY_POWER_LINE)
        types = np.asarray([k for p, k in zip(parts, kinds) for _ in range(len(p))])
        return (np.concatenate(parts) if concatenated_output else parts), types

    def set_substation_switches_in_action(self, action, substation_id, new_values):
        action = self.array_to_action(action)
        new_values = np.asarray(new_values)
        _, types = self.get_substation_switches_in_action(action, substation_id, concatenated_output=False)
        assert len(types) == len(new_values), 'Expected new_values of size %d for substation %d, got size %d' % (
            len(types), substation_id, len(new_values))
        action.prods_switches_subaction[self.prods_subs_ids == substation_id] = new_values[
            types == ElementType.PRODUCTION]
        action.loads_switches_subaction[self.loads_subs_ids == substation_id] = new_values[
            types == ElementType.CONSUMPTION]
        action.lines_or_switches_subaction[self.lines_or_subs_id == substation_id] = new_values[
            types == ElementType.ORIGIN_POWER_LINE]
        action.lines_ex_switches_subaction[self.lines_ex_subs_id == substation_id] = new_values[
            types == ElementType.EXTREMITY_POWER_LINE]
        return action

    def get_lines_status_switches_of_substation(self, action, substation_id):
        assert substation_id in self.substations_ids, 'Substation with id %d does not exist' % substation_id
        return action.lines_status_subaction[(self.lines_or_subs_id == substation_id) |
                                             (self.lines_ex_subs_id == substation_id)]

    @staticmethod
    def get_lines_status_switch_from_id(action, line_id):
        return action.lines_status_subaction[line_id]

    @staticmethod
    def set_lines_status_switch_from_id(action, line_id, new_switch_value):
        action.lines_status_subaction[line_id] = new_switch_value


# ---------------------------------------------------------------------------------------------------- observation
def observation_fields(G, L, N, S):
    """(name, size) in Observation.as_array order (environment.py:451-466, 511-517, 583-595)."""
    minimalist = [('active_loads', L), ('are_loads_cut', L), ('planned_active_loads', L), ('loads_nodes', L),
                  ('active_productions', G), ('are_productions_cut', G), ('planned_active_productions', G),
                  ('productions_nodes', G), ('lines_or_nodes', N), ('lines_ex_nodes', N), ('ampere_flows', N),
                  ('lines_status', N), ('timesteps_before_lines_reconnectable', N),
                  ('timesteps_before_lines_reactionable', N), ('timesteps_before_nodes_reactionable', S),
                  ('timesteps_before_planned_maintenance', N), ('date_year', 1), ('date_month', 1), ('date_day', 1),
                  ('date_hour', 1), ('date_minute', 1), ('date_second', 1)]
    ac = [('reactive_loads', L), ('voltage_loads', L), ('reactive_productions', G), ('voltage_productions', G),
          ('active_flows_origin', N), ('reactive_flows_origin', N), ('voltage_flows_origin', N),
          ('active_flows_extremity', N), ('reactive_flows_extremity', N), ('voltage_flows_extremity', N),
          ('planned_reactive_loads', L), ('planned_voltage_productions', G)]
    static = [('substations_ids', S), ('loads_substations_ids', L), ('productions_substations_ids', G),
              ('lines_or_substations_ids', N), ('lines_ex_substations_ids', N), ('thermal_limits', N),
              ('initial_productions_nodes', G), ('initial_loads_nodes', L), ('initial_lines_or_nodes', N),
              ('initial_lines_ex_nodes', N)]
    return minimalist, ac, static


class Observation(object):
    """All fields of the reference's Observation as attributes; as_array() is the flat vector agents consume."""

    def __init__(self, **fields):
        self._order = list(fields.keys())
        for k, v in fields.items():
            setattr(self, k, v)

    def as_array(self):
        return np.concatenate([np.atleast_1d(np.asarray(getattr(self, k), dtype=np.float64)).ravel()
                               for k in self._order])

    def as_dict(self):
        return OrderedDict((k, getattr(self, k)) for k in self._order)

    def get_nodes_of_substation(self, substation_id):
        """(node values, element types) of the elements of a substation: prods, loads, line origins, line
        extremities (environment.py:604-640)."""
        assert substation_id in self.substations_ids, 'Substation with id %d does not exist' % substation_id
        parts = (self.productions_nodes[self.productions_substations_ids == substation_id],
                 self.loads_nodes[self.loads_substations_ids == substation_id],
                 self.lines_or_nodes[self.lines_or_substations_ids == substation_id],
                 self.lines_ex_nodes[self.lines_ex_substations_ids == substation_id])
        kinds = (ElementType.PRODUCTION, ElementType.CONSUMPTION, ElementType.ORIGIN_POWER_LINE,
                 ElementType.EXTREMITY_POWER_LINE)
        return np.concatenate(parts), [k for p, k in zip(parts, kinds) for _ in range(len(p))]

    def get_lines_capacity_usage(self):
        return np.asarray(self.ampere_flows) / np.asarray(self.thermal_limits)


class ObservationSpace(object):
    """environment.py:277-403 without gym: knows the flat layout and converts arrays back to Observation."""

    def __init__(self, number_generators, number_consumers, number_power_lines, number_substations,
                 n_timesteps_horizon_maintenance):
        self.number_productions = number_generators
        self.number_loads = number_consumers
        self.number_power_lines = number_power_lines
        self.number_substations = number_substations
        self.n_timesteps_horizon_maintenance = n_timesteps_horizon_maintenance
        self.grid_number_of_elements = number_generators + number_consumers + 2 * number_power_lines
        m, ac, st = observation_fields(number_generators, number_consumers, number_power_lines, number_substations)
        self._fields = m + ac + st
        self.shape = tuple((n,) for _, n in self._fields)
        self.length = sum(n for _, n in self._fields)
        self.dynamic_length = sum(n for _, n in m + ac)

    def array_to_observation(self, array):
        if len(array) != self.length:
            raise ValueError('Expected observation array of length %d, got %d' % (self.length, len(array)))
        array = np.asarray(array)
        out, off = OrderedDict(), 0
        for name, n in self._fields:
            out[name] = array[off:off + n]
            off += n
        return Observation(**out)


# --------------------------------------------------------------------------------------------------------- RunEnv
class _GameView(object):
    """The few `env.game.*` accessors agents and the Runner use (game.py:342-403, 980-1100)."""

    def __init__(self, env):
        self._env = env
        case = env._vec.case
        self.n_timesteps_horizon_maintenance = int(env._vec.config['n_timesteps_horizon_maintenance'])
        self.substations_ids = case.sub_ids
        self.epoch = 1

    def get_number_elements(self):
        c = self._env._vec.case
        return c.n_gen, c.n_load, c.n_line, c.n_sub

    def get_substations_ids(self):
        return self._env._vec.case.sub_ids

    def get_substations_ids_prods(self):
        c = self._env._vec.case
        return c.sub_ids[c.gen_sub]

    def get_substations_ids_loads(self):
        c = self._env._vec.case
        return c.sub_ids[c.load_sub]

    def get_substations_ids_lines_or(self):
        c = self._env._vec.case
        return c.sub_ids[c.line_or_sub]

    def get_substations_ids_lines_ex(self):
        c = self._env._vec.case
        return c.sub_ids[c.line_ex_sub]

    def get_current_chronic_name(self):
        return self._env.get_current_chronic_name()

    def get_current_datetime(self):
        return self._env.get_current_datetime()

    def is_action_valid(self, action):
        return self._env.is_action_valid(action)

    def get_reward_signal_class(self):
        return self._env.reward_signal


class RunEnv(object):
    """Drop-in for pypownet.environment.RunEnv (environment.py:788-914), one env on one GPU."""

    def __init__(self, parameters_folder, game_level, chronic_looping_mode='natural', start_id=0,
                 game_over_mode='soft', renderer_latency=None, without_overflow_cutoff=False, seed=None, device=0):
        self.parameters_folder = parameters_folder
        self.game_level = game_level
        self.chronic_looping_mode = chronic_looping_mode
        self.start_id = start_id
        self.game_over_mode = game_over_mode
        self.renderer_latency = renderer_latency
        self.without_overflow_cutoff = without_overflow_cutoff
        self.device = device
        self.seed = seed
        self.game = None
        self.action_space = None
        self.observation_space = None
        self.reward_signal = None
        self.last_rewards = None
        self._vec = None
        if seed is not None:
            np.random.seed(seed)
        self.reset()

    def reset(self):
        from pypownet_b200.vec_env import VecRunEnv
        if self._vec is not None:
            self._vec.close()
        self._vec = VecRunEnv.from_folder(self.parameters_folder, self.game_level, n_envs=1,
                                          chronic_looping_mode=self.chronic_looping_mode, start_id=self.start_id,
                                          game_over_mode=self.game_over_mode,
                                          without_overflow_cutoff=self.without_overflow_cutoff, device=self.device,
                                          seed=self.seed or 0)
        case = self._vec.case
        self.game = _GameView(self)
        ids = case.sub_ids
        self.action_space = ActionSpace(case.n_gen, case.n_load, case.n_line, case.n_sub, substations_ids=ids,
                                        prods_subs_ids=ids[case.gen_sub], loads_subs_ids=ids[case.load_sub],
                                        lines_or_subs_id=ids[case.line_or_sub], lines_ex_subs_id=ids[case.line_ex_sub])
        self.observation_space = ObservationSpace(case.n_gen, case.n_load, case.n_line, case.n_sub,
                                                  self.game.n_timesteps_horizon_maintenance)
        self.reward_signal = self._vec.parameters.get_reward_signal_class()()
        self._custom_reward = not self._is_shipped_reward(self.reward_signal)
        self.last_rewards = []
        self._last_obs = self._vec.obs[0].cpu().numpy().copy()
        return self.get_observation(True)

    @staticmethod
    def _is_shipped_reward(sig):
        """True when the plug-in is the shipped five-term reward WITH the shipped constants (parameters/default14/
        reward_signal.py:11-43: every factor is a fixed multiple of `constant`): only then does the step kernel's reward
        equal the plug-in's.  A plug-in that keeps the attribute names but tunes a factor is evaluated on the host."""
        k = getattr(sig, 'too_many_productions_cut', None)
        if k is None:
            return False
        c = -float(k)
        expect = {'multiplicative_factor_line_usage_reward': -1., 'multiplicative_factor_distance_initial_grid': -.02,
                  'multiplicative_factor_number_loads_cut': -c / 5., 'multiplicative_factor_number_prods_cut': -c / 10.,
                  'connexity_exception_reward': -c, 'loadflow_exception_reward': -c,
                  'multiplicative_factor_number_illegal_broken_line_switch': -c / 100.,
                  'multiplicative_factor_number_illegal_oncooldown_line_switch': -c / 100.,
                  'multiplicative_factor_number_illegal_oncooldown_substation_switch': -c / 100.,
                  'multiplicative_factor_number_illegal_lines_reconnection': -c / 100.,
                  'too_many_consumptions_cut': -c, 'too_much_activated_elements': -5 * c,
                  'multiplicative_factor_number_line_switches': -.2, 'multiplicative_factor_number_node_switches': -.1}
        for name, value in expect.items():
            if hasattr(sig, name) and float(getattr(sig, name)) != value:
                return False
        return True

    def _corrected_action(self, action, illegal_row):
        """The action as the reference leaves it after an IllegalActionException (game.py:809-854 mutates the submitted
        action in place): everything dropped when too many elements were activated, else the illegal line switches and
        the switches of substations on cooldown zeroed."""
        c = self._vec.case
        N, S = c.n_line, c.n_sub
        a = np.array(action.as_array(), dtype=np.int64)
        nt = c.n_gen + c.n_load + 2 * N
        if illegal_row[0]:
            a[:] = 0
        else:
            bad_lines = illegal_row[1:1 + N].astype(bool) | illegal_row[1 + N:1 + 2 * N].astype(bool)
            a[nt:][bad_lines] = 0
            a[:nt][illegal_row[1 + 2 * N:].astype(bool)[np.asarray(c.elem_sub)]] = 0
        return self.action_space.array_to_action(a)

    def get_observation(self, as_array=True):
        return self._last_obs.copy() if as_array else self.observation_space.array_to_observation(self._last_obs)

    def _get_obs(self):
        return self.get_observation(False)

    def is_action_valid(self, action):
        a = self.action_space._verify_action_shape(action).as_array().astype(np.uint8)[None]
        return bool(self._vec.is_action_valid(a)[0].item())

    def _flag_object(self, code, illegal_row):
        N, S = self._vec.case.n_line, self._vec.case.n_sub
        if code == _lib.FLAG_NONE:
            return None
        if code == _lib.FLAG_DIVERGING_LOADFLOW:
            return DivergingLoadflowException(None, 'The grid is not connexe or the loadflow has diverged')
        if code == _lib.FLAG_TOO_MANY_LOADS_CUT:
            return TooManyConsumptionsCut('There are too many isolated loads')
        if code == _lib.FLAG_TOO_MANY_PRODS_CUT:
            return TooManyProductionsCut('There are too many isolated productions')
        return IllegalActionException('Some switches of the action were illegal and have been ignored',
                                      bool(illegal_row[0]), illegal_row[1:1 + N].astype(bool),
                                      illegal_row[1 + N:1 + 2 * N].astype(bool), illegal_row[1 + 2 * N:].astype(bool))

    def _finish(self, obs_row, reward_row, done, code, illegal_row, action, do_sum):
        flag = self._flag_object(code, illegal_row)
        observation = None if done else self.observation_space.array_to_observation(obs_row)
        if self._custom_reward:                      # user plug-in: evaluated on the host, as the reference does
            if code == _lib.FLAG_ILLEGAL_ACTION or illegal_row.any():
                action = self._corrected_action(action, illegal_row)
            reward_aslist = self.reward_signal.compute_reward(observation=observation, action=action, flag=flag)
        else:                                        # the shipped five-term reward is computed by the step kernel
            reward_aslist = [float(v) for v in reward_row]
        self.last_rewards = reward_aslist
        return (None if done else obs_row), (sum(reward_aslist) if do_sum else reward_aslist), bool(done), flag

    def step(self, action, do_sum=True):
        submitted = self.action_space._verify_action_shape(action)
        a = submitted.as_array().astype(np.uint8)[None]
        obs, reward, done, flag = self._vec.step(a)
        d = bool(done[0].item())
        row = obs[0].cpu().numpy().copy()
        if not d:
            self._last_obs = row
        return self._finish(row, reward[0].cpu().numpy(), d, int(flag[0].item()),
                            self._vec.illegal[0].cpu().numpy(), submitted, do_sum)

    def simulate(self, action, do_sum=True):
        to_simulate = self.action_space._verify_action_shape(action)
        a = to_simulate.as_array().astype(np.uint8)[None]
        obs, reward, done, flag = self._vec.simulate(a)
        ill = self._vec.sim_illegal[0].cpu().numpy()
        return self._finish(obs[0].cpu().numpy(), reward[0].cpu().numpy(), bool(done[0].item()),
                            int(flag[0].item()), ill, to_simulate, do_sum)

    def process_game_over(self):
        obs = self._vec.process_game_over()
        self._last_obs = obs[0].cpu().numpy().copy()
        self.game.epoch += 1
        return self.get_observation()

    def render(self, game_over=False):
        raise NotImplementedError('the pygame renderer of the reference is out of scope (SURVEY.md section 2, row 12)')

    def get_current_chronic_name(self):
        cur = self._vec.get_state(_lib.STATE_COUNTERS)[0].cpu().numpy()
        return self._vec.chronics[int(cur[-4])].name

    def get_current_datetime(self):
        from datetime import datetime
        o = self._last_obs
        c = self._vec.case
        off = 4 * c.n_load + 4 * c.n_gen + 7 * c.n_line + c.n_sub
        y, mo, d, h, mi, s = [int(v) for v in o[off:off + 6]]
        return datetime(y, mo, d, h, mi, s)


OBSERVATION_MEANING = {
    'active_loads': 'the active power of the loads (MW)', 'are_loads_cut': 'mask of isolated loads',
    'planned_active_loads': 'the active power of the loads planned for the next timestep (MW)',
    'loads_nodes': 'node (0 or 1) of each load within its substation',
    'active_productions': 'the active power of the productions (MW)', 'are_productions_cut': 'mask of isolated productions',
    'ampere_flows': 'current in each line (A)', 'lines_status': 'mask of in-service lines',
    'thermal_limits': 'maximum current of each line (A)',
}
