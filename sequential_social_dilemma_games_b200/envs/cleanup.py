"""CleanupEnv with the reference's interface (social_dilemmas/envs/cleanup.py), CUDA-backed."""
import random

import numpy as np

from ..config import (APPLE_RESPAWN_PROBABILITY, BEAM_LENGTH, CLEANUP_COLOURS, DEFAULT_VIEW_SIZE, KIND_CLEANUP,
                      THRESHOLD_DEPLETION, THRESHOLD_RESTORATION, WASTE_SPAWN_PROBABILITY, cleanup_probabilities)
from ..maps import CLEANUP_MAP
from ._game_env import GameEnv
from .agent import CleanupAgent
from .map_env import ACTIONS

# module-level knobs under the reference's names (cleanup.py:11-27)
ACTIONS['FIRE'] = ACTIONS['CLEAN'] = BEAM_LENGTH
CLEANUP_COLORS = {ch: list(rgb) for ch, rgb in CLEANUP_COLOURS.items()}
CLEANUP_VIEW_SIZE = DEFAULT_VIEW_SIZE
thresholdDepletion, thresholdRestoration = THRESHOLD_DEPLETION, THRESHOLD_RESTORATION
wasteSpawnProbability, appleRespawnProbability = WASTE_SPAWN_PROBABILITY, APPLE_RESPAWN_PROBABILITY


class CleanupEnv(GameEnv):
    KIND = KIND_CLEANUP
    VIEW_SIZE = CLEANUP_VIEW_SIZE
    AGENT_CLASS = CleanupAgent
    NUM_ACTIONS = 9

    def __init__(self, ascii_map=CLEANUP_MAP, num_agents=1, render=False, return_agent_actions=False, device="cuda:0"):
        super().__init__(ascii_map, num_agents, render, return_agent_actions=return_agent_actions, device=device)
        # cleanup.py:36-41: the area that can hold waste, and a first evaluation of the probabilities on the still blank
        # world_map (MapEnv.__init__ has only placed the agents so far)
        self.potential_waste_area = int(np.count_nonzero(np.isin(self.base_map, ('H', 'R'))))
        self.current_apple_spawn_prob, self.current_waste_spawn_prob = appleRespawnProbability, wasteSpawnProbability
        self.compute_probabilities()
        # cleanup.py:43-62, row-major; the spawn points are listed a second time, as upstream does
        self.spawn_points += self.cells('P')
        self.apple_points = self.cells('B')
        self.stream_points = self.cells('S')
        self.river_points = self.cells('R')
        self.waste_start_points = self.cells('H')
        self.waste_points = self.cells('H', 'R')
        self.color_map.update(CLEANUP_COLORS)   # cleanup.py:64: this mutates the shared default table, as upstream does
        self.view_len = CLEANUP_VIEW_SIZE

    def _config_kwargs(self):
        return dict(beam_length=ACTIONS['FIRE'],
                    cleanup_params=dict(threshold_depletion=thresholdDepletion, threshold_restoration=thresholdRestoration,
                                        waste_spawn_probability=wasteSpawnProbability,
                                        apple_respawn_probability=appleRespawnProbability))

    # ------------------------------------------------------------------ hooks
    def custom_reset(self):
        """cleanup.py:84-92: waste, river and stream back where the base map has them."""
        for char, points in (('H', self.waste_start_points), ('R', self.river_points), ('S', self.stream_points)):
            for row, col in points:
                self.world_map[row, col] = char
        self.compute_probabilities()

    def custom_map_update(self):
        self.compute_probabilities()
        self.update_map(self.spawn_apples_and_waste())

    def spawn_apples_and_waste(self):
        """cleanup.py:132-154 on the device.  random.shuffle(self.waste_points) is called here, as in
        the reference, exactly when the waste probability is non-zero; the shuffled order and the
        np.random.rand stream are handed to the kernel."""
        if np.isclose(self.current_waste_spawn_prob, 0):
            return self._device_spawn(waste_order=None)
        random.shuffle(self.waste_points)
        width = self.world_map.shape[1]
        return self._device_spawn(waste_order=np.array([[r * width + c for r, c in self.waste_points]], dtype=np.uint16))

    def compute_probabilities(self):
        """cleanup.py:156-171 -- config.cleanup_probabilities is the same expression, the one the device tables are
        built from."""
        self.current_apple_spawn_prob, self.current_waste_spawn_prob = cleanup_probabilities(
            self.potential_waste_area - self.compute_permitted_area(), self.potential_waste_area,
            thresholdDepletion, thresholdRestoration, wasteSpawnProbability, appleRespawnProbability)

    def compute_permitted_area(self):
        """cleanup.py:173-179: waste area not covered by waste right now."""
        return self.potential_waste_area - int(np.count_nonzero(self.world_map == 'H'))
