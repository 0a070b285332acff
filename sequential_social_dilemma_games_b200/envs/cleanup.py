"""CleanupEnv with the reference's interface (social_dilemmas/envs/cleanup.py), CUDA-backed."""
import random

import numpy as np

from ..config import (APPLE_RESPAWN_PROBABILITY, BEAM_LENGTH, CLEANUP_COLOURS, DEFAULT_VIEW_SIZE, KIND_CLEANUP,
                      THRESHOLD_DEPLETION, THRESHOLD_RESTORATION, WASTE_SPAWN_PROBABILITY)
from ..maps import CLEANUP_MAP
from .agent import CleanupAgent
from .map_env import ACTIONS, MapEnv
from .spaces import Box, Dict, Discrete

ACTIONS['FIRE'] = BEAM_LENGTH
ACTIONS['CLEAN'] = BEAM_LENGTH
CLEANUP_COLORS = {k: list(v) for k, v in CLEANUP_COLOURS.items()}   # cleanup.py:15-18
CLEANUP_VIEW_SIZE = DEFAULT_VIEW_SIZE                               # cleanup.py:22
thresholdDepletion = THRESHOLD_DEPLETION                            # cleanup.py:24-27
thresholdRestoration = THRESHOLD_RESTORATION
wasteSpawnProbability = WASTE_SPAWN_PROBABILITY
appleRespawnProbability = APPLE_RESPAWN_PROBABILITY


class CleanupEnv(MapEnv):
    KIND = KIND_CLEANUP
    VIEW_SIZE = CLEANUP_VIEW_SIZE

    def __init__(self, ascii_map=CLEANUP_MAP, num_agents=1, render=False, return_agent_actions=False, device="cuda:0"):
        super().__init__(ascii_map, num_agents, render, return_agent_actions=return_agent_actions, device=device)
        unique, counts = np.unique(self.base_map, return_counts=True)
        counts_dict = dict(zip(unique, counts))
        self.potential_waste_area = counts_dict.get('H', 0) + counts_dict.get('R', 0)
        self.current_apple_spawn_prob = appleRespawnProbability
        self.current_waste_spawn_prob = wasteSpawnProbability
        self.compute_probabilities()
        self.apple_points, self.waste_start_points, self.waste_points = [], [], []
        self.river_points, self.stream_points = [], []
        for row in range(self.base_map.shape[0]):
            for col in range(self.base_map.shape[1]):
                ch = self.base_map[row, col]
                if ch == 'P':
                    self.spawn_points.append([row, col])   # cleanup.py:51-52: every 'P' a second time
                elif ch == 'B':
                    self.apple_points.append([row, col])
                elif ch == 'S':
                    self.stream_points.append([row, col])
                if ch == 'H':
                    self.waste_start_points.append([row, col])
                if ch == 'H' or ch == 'R':
                    self.waste_points.append([row, col])
                if ch == 'R':
                    self.river_points.append([row, col])
        self.color_map.update(CLEANUP_COLORS)   # cleanup.py:64 (mutates the shared default table, as upstream)
        self.view_len = CLEANUP_VIEW_SIZE

    def _config_kwargs(self):
        return dict(beam_length=ACTIONS['FIRE'],
                    cleanup_params=dict(threshold_depletion=thresholdDepletion, threshold_restoration=thresholdRestoration,
                                        waste_spawn_probability=wasteSpawnProbability,
                                        apple_respawn_probability=appleRespawnProbability))

    @property
    def action_space(self):
        return Discrete(9)

    @property
    def observation_space(self):
        v = 2 * self.view_len + 1
        if self.return_agent_actions:
            return Dict({"curr_obs": Box(low=-np.inf, high=np.inf, shape=(v, v, 3), dtype=np.float32),
                         "other_agent_actions": Box(low=0, high=len(ACTIONS), shape=(self.num_agents - 1,), dtype=np.int32),
                         "visible_agents": Box(low=0, high=self.num_agents, shape=(self.num_agents - 1,), dtype=np.int32)})
        return Box(low=0.0, high=0.0, shape=(v, v, 3), dtype=np.float32)

    def custom_reset(self):
        for p in self.waste_start_points:
            self.world_map[p[0], p[1]] = 'H'
        for p in self.river_points:
            self.world_map[p[0], p[1]] = 'R'
        for p in self.stream_points:
            self.world_map[p[0], p[1]] = 'S'
        self.compute_probabilities()

    def custom_map_update(self):
        self.compute_probabilities()
        self.update_map(self.spawn_apples_and_waste())

    def setup_agents(self):
        map_with_agents = self.get_map_with_agents()
        for i in range(self.num_agents):
            agent_id = 'agent-' + str(i)
            spawn_point = self.spawn_point()
            rotation = self.spawn_rotation()
            self.agents[agent_id] = CleanupAgent(agent_id, spawn_point, rotation, map_with_agents, view_len=CLEANUP_VIEW_SIZE)

    def spawn_apples_and_waste(self):
        """cleanup.py:132-154 on the device.  random.shuffle(self.waste_points) is called here, as in
        the reference, exactly when the waste probability is non-zero; the shuffled order and the
        np.random.rand stream are handed to the kernel."""
        waste_order = None
        if not np.isclose(self.current_waste_spawn_prob, 0):
            random.shuffle(self.waste_points)
            w = self.world_map.shape[1]
            waste_order = np.array([[r * w + c for r, c in self.waste_points]], dtype=np.uint16)
        return self._device_spawn(waste_order=waste_order)

    def compute_probabilities(self):
        """cleanup.py:156-171 (the device evaluates the same expression through the table config.py builds)."""
        waste_density = 0
        if self.potential_waste_area > 0:
            waste_density = 1 - self.compute_permitted_area() / self.potential_waste_area
        if waste_density >= thresholdDepletion:
            self.current_apple_spawn_prob = 0
            self.current_waste_spawn_prob = 0
        else:
            self.current_waste_spawn_prob = wasteSpawnProbability
            if waste_density <= thresholdRestoration:
                self.current_apple_spawn_prob = appleRespawnProbability
            else:
                self.current_apple_spawn_prob = (1 - (waste_density - thresholdRestoration)
                                                 / (thresholdDepletion - thresholdRestoration)) * appleRespawnProbability

    def compute_permitted_area(self):
        unique, counts = np.unique(self.world_map, return_counts=True)
        current_area = dict(zip(unique, counts)).get('H', 0)
        return self.potential_waste_area - current_area
