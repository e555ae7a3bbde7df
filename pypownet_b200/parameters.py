"""Environment folder layout and configuration.yaml access (pypownet/parameters.py:35-153)."""
import importlib.util
import logging
import os

import yaml

from pypownet_b200.reward_signal import RewardSignal

MANDATORY_KEYS = (
    'loadflow_backend', 'loadflow_mode', 'max_seconds_per_timestep', 'hard_overflow_coefficient',
    'n_timesteps_hard_overflow_is_broken', 'n_timesteps_consecutive_soft_overflow_breaks',
    'n_timesteps_soft_overflow_is_broken', 'n_timesteps_horizon_maintenance', 'max_number_prods_game_over',
    'max_number_loads_game_over', 'n_timesteps_actionned_line_reactionable',
    'n_timesteps_actionned_node_reactionable', 'max_number_actionned_substations', 'max_number_actionned_lines',
    'max_number_actionned_total')


class Parameters(object):
    def __init__(self, parameters_folder, game_level):
        self.logger = logging.getLogger('pypownet.' + __name__)
        self.parameters_path = os.path.abspath(parameters_folder)
        if not os.path.exists(self.parameters_path):
            raise FileNotFoundError('folder %s does not exist' % self.parameters_path)
        self.level_folder = os.path.join(self.parameters_path, game_level)
        if not os.path.exists(self.level_folder):
            raise FileNotFoundError('Game level folder %s does not exist in %s' % (game_level, self.parameters_path))
        for f in ('configuration.yaml', 'reference_grid.py', 'chronics'):
            if not os.path.exists(os.path.join(self.level_folder, f)):
                raise FileNotFoundError('Mandatory file/folder %s not found within %s' % (f, self.level_folder))
        self.reference_grid_path = os.path.join(self.level_folder, 'reference_grid.py')
        self.chronics_path = os.path.join(self.level_folder, 'chronics')
        with open(os.path.join(self.level_folder, 'configuration.yaml')) as stream:
            self.simulator_configuration = yaml.safe_load(stream)
        missing = [k for k in MANDATORY_KEYS if k not in self.simulator_configuration]
        if missing:
            raise KeyError('configuration.yaml lacks %s' % ', '.join(missing))
        backend = str(self.simulator_configuration['loadflow_backend']).lower()
        if backend not in ('matpower', 'pypower'):
            raise ValueError('loadflow_backend %s is not currently supported' % backend)
        mode = str(self.simulator_configuration['loadflow_mode']).lower()
        if mode not in ('ac', 'dc'):
            raise ValueError('loadflow_mode value in configuration file should be either "AC" or "DC"')
        # custom reward plug-in: <parameters_folder>/reward_signal.py with class CustomRewardSignal (:57-70)
        self.reward_signal_class = RewardSignal
        path = os.path.join(self.parameters_path, 'reward_signal.py')
        if os.path.exists(path):
            try:
                self.reward_signal_class = _load_reward_class(path)
            except ImportError:
                self.logger.error('/!\\ Using default reward signal, reward_signal.py could not be imported')
        else:
            self.logger.error('/!\\ Using default reward signal, as reward_signal.py file is not found')

    def get_reward_signal_class(self):
        return self.reward_signal_class

    def get_reference_grid_path(self, loadflow_backend='pypower'):
        return self.reference_grid_path

    def get_chronics_path(self):
        return self.chronics_path

    def get_parameters_path(self):
        return self.parameters_path

    def get_loadflow_backend(self):
        return str(self.simulator_configuration['loadflow_backend']).lower()

    def is_dc_mode(self):
        return str(self.simulator_configuration['loadflow_mode']).lower() == 'dc'

    def __getattr__(self, name):
        # get_<key>() accessors for every configuration key, as the reference exposes (:110-153)
        if name.startswith('get_') and name[4:] in self.__dict__.get('simulator_configuration', {}):
            key = name[4:]
            return lambda: self.simulator_configuration[key]
        raise AttributeError(name)

    def __str__(self):
        params_str = ['    ' + k + ': ' + str(v) for k, v in self.simulator_configuration.items()]
        width = max(map(len, params_str))
        return '\n'.join(['  ' + '=' * width, ' ' * (width // 2 - 5) + 'GAME PARAMETERS', '  ' + '=' * width,
                          '\n'.join(params_str), '  ' + '=' * width])


def _load_reward_class(path):
    """Import CustomRewardSignal from an environment's reward_signal.py.  Those files are written against the
    reference's module names (`import pypownet.environment`, `pypownet.reward_signal`); alias them to this package
    so an unmodified environment folder plugs in."""
    import sys
    import pypownet_b200
    import pypownet_b200.environment
    import pypownet_b200.reward_signal
    if 'pypownet' not in sys.modules:
        sys.modules['pypownet'] = pypownet_b200
        sys.modules['pypownet.environment'] = pypownet_b200.environment
        sys.modules['pypownet.reward_signal'] = pypownet_b200.reward_signal
        sys.modules['pypownet.game'] = pypownet_b200.environment
    spec = importlib.util.spec_from_file_location('reward_signal_%x' % (hash(path) & 0xffffffff), path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return getattr(mod, 'CustomRewardSignal')
