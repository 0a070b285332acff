"""HarvestEnv with the reference's interface (social_dilemmas/envs/harvest.py), CUDA-backed."""
import numpy as np

from ..config import BEAM_LENGTH, DEFAULT_VIEW_SIZE, HARVEST_SPAWN_PROB, KIND_HARVEST
from ..maps import HARVEST_MAP
from .agent import HarvestAgent
from .map_env import ACTIONS, MapEnv
from .spaces import Box, Dict, Discrete

APPLE_RADIUS = 2                        # harvest.py:8 (j*j + k*k <= 2: the 3x3 window)
SPAWN_PROB = list(HARVEST_SPAWN_PROB)   # harvest.py:13
HARVEST_VIEW_SIZE = DEFAULT_VIEW_SIZE   # harvest.py:15
ACTIONS['FIRE'] = BEAM_LENGTH           # harvest.py:11


class HarvestEnv(MapEnv):
    KIND = KIND_HARVEST
    VIEW_SIZE = HARVEST_VIEW_SIZE

    def __init__(self, ascii_map=HARVEST_MAP, num_agents=1, render=False, return_agent_actions=False, device="cuda:0"):
        super().__init__(ascii_map, num_agents, render, return_agent_actions=return_agent_actions, device=device)
        self.apple_points = []
        for row in range(self.base_map.shape[0]):
            for col in range(self.base_map.shape[1]):
                if self.base_map[row, col] == 'A':
                    self.apple_points.append([row, col])
        self.view_len = HARVEST_VIEW_SIZE

    def _config_kwargs(self):
        return dict(harvest_spawn_prob=SPAWN_PROB, beam_length=ACTIONS['FIRE'])

    @property
    def observation_space(self):
        v = 2 * self.view_len + 1
        if self.return_agent_actions:
            return Dict({"curr_obs": Box(low=-np.inf, high=np.inf, shape=(v, v, 3), dtype=np.float32),
                         "other_agent_actions": Box(low=0, high=len(ACTIONS), shape=(self.num_agents - 1,), dtype=np.int32),
                         "visible_agents": Box(low=0, high=self.num_agents, shape=(self.num_agents - 1,), dtype=np.int32)})
        return Box(low=0.0, high=0.0, shape=(v, v, 3), dtype=np.float32)   # harvest.py:39-40

    @property
    def action_space(self):
        return Discrete(8)

    def setup_agents(self):
        """harvest.py:46-55."""
        map_with_agents = self.get_map_with_agents()
        for i in range(self.num_agents):
            agent_id = 'agent-' + str(i)
            spawn_point = self.spawn_point()
            rotation = self.spawn_rotation()
            self.agents[agent_id] = HarvestAgent(agent_id, spawn_point, rotation, map_with_agents, view_len=HARVEST_VIEW_SIZE)

    def custom_reset(self):
        for apple_point in self.apple_points:
            self.world_map[apple_point[0], apple_point[1]] = 'A'

    def custom_map_update(self):
        self.update_map(self.spawn_apples())

    def spawn_apples(self):
        """harvest.py:75-104 on the device: one np.random.rand per eligible apple point, in row-major
        order; returns [(row, col, 'A')] of the points that spawn."""
        return self._device_spawn()

    def count_apples(self, window):
        unique, counts = np.unique(window, return_counts=True)
        return dict(zip(unique, counts)).get('A', 0)
