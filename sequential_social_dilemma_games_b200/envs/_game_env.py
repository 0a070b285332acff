"""What HarvestEnv and CleanupEnv share on top of MapEnv: spaces, agent construction, cell bookkeeping."""
import numpy as np

from .map_env import ACTIONS, MapEnv
from .spaces import Box, Dict, Discrete


class GameEnv(MapEnv):
    AGENT_CLASS = None      # HarvestAgent / CleanupAgent
    NUM_ACTIONS = 0         # 8 (harvest.py:42-44) / 9 (cleanup.py:68-70)

    def cells(self, *chars):
        """[row, col] lists of the base-map cells holding one of `chars`, row-major (the order every per-cell loop of the
        reference runs in: harvest.py:22-26, cleanup.py:43-62)."""
        rows, cols = np.nonzero(np.isin(self.base_map, chars))
        return [[int(r), int(c)] for r, c in zip(rows, cols)]

    @property
    def action_space(self):
        return Discrete(self.NUM_ACTIONS)

    @property
    def observation_space(self):
        """harvest.py:30-40 / cleanup.py:72-82: a (2v+1)^2 RGB box, or the MOA dict with the other agents' actions."""
        side = 2 * self.view_len + 1
        if not self.return_agent_actions:
            return Box(low=0.0, high=0.0, shape=(side, side, 3), dtype=np.float32)
        others = (self.num_agents - 1,)
        return Dict({"curr_obs": Box(low=-np.inf, high=np.inf, shape=(side, side, 3), dtype=np.float32),
                     "other_agent_actions": Box(low=0, high=len(ACTIONS), shape=others, dtype=np.int32),
                     "visible_agents": Box(low=0, high=self.num_agents, shape=others, dtype=np.int32)})

    def setup_agents(self):
        """harvest.py:46-55 / cleanup.py:118-130: one spawn_point() and one spawn_rotation() per agent, in id order, all
        agents sharing the overlay grid taken before the first of them exists."""
        overlay = self.get_map_with_agents()
        for index in range(self.num_agents):
            name = 'agent-%d' % index
            where, facing = self.spawn_point(), self.spawn_rotation()
            self.agents[name] = self.AGENT_CLASS(name, where, facing, overlay, view_len=self.VIEW_SIZE)
