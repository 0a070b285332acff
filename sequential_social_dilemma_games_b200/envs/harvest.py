"""HarvestEnv with the reference's interface (social_dilemmas/envs/harvest.py), CUDA-backed."""
import numpy as np

from ..config import BEAM_LENGTH, DEFAULT_VIEW_SIZE, HARVEST_SPAWN_PROB, KIND_HARVEST
from ..maps import HARVEST_MAP
from ._game_env import GameEnv
from .agent import HarvestAgent
from .map_env import ACTIONS

# module-level knobs under the reference's names (harvest.py:8-15)
APPLE_RADIUS = 2                        # j*j + k*k <= 2: the 3x3 window
SPAWN_PROB = list(HARVEST_SPAWN_PROB)
HARVEST_VIEW_SIZE = DEFAULT_VIEW_SIZE
ACTIONS['FIRE'] = BEAM_LENGTH


class HarvestEnv(GameEnv):
    KIND = KIND_HARVEST
    VIEW_SIZE = HARVEST_VIEW_SIZE
    AGENT_CLASS = HarvestAgent
    NUM_ACTIONS = 8

    def __init__(self, ascii_map=HARVEST_MAP, num_agents=1, render=False, return_agent_actions=False, device="cuda:0"):
        super().__init__(ascii_map, num_agents, render, return_agent_actions=return_agent_actions, device=device)
        self.apple_points = self.cells('A')   # harvest.py:22-26, row-major
        self.view_len = HARVEST_VIEW_SIZE

    def _config_kwargs(self):
        return dict(harvest_spawn_prob=SPAWN_PROB, beam_length=ACTIONS['FIRE'])

    # ------------------------------------------------------------------ hooks
    def custom_reset(self):
        """harvest.py:57-60: every apple point starts with an apple."""
        for row, col in self.apple_points:
            self.world_map[row, col] = 'A'

    def custom_map_update(self):
        self.update_map(self.spawn_apples())

    def spawn_apples(self):
        """harvest.py:75-104 on the device: one np.random.rand per eligible apple point, in row-major
        order; returns [(row, col, 'A')] of the points that spawn."""
        return self._device_spawn()

    def count_apples(self, window):
        """harvest.py:106-114."""
        return int(np.count_nonzero(np.asarray(window) == 'A'))
