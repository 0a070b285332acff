"""Host-side agent proxies with the reference's interface (social_dilemmas/envs/agent.py).

In the batched engine an agent is three bytes of device state; these objects exist so that code
written against the reference -- `env.agents['agent-0'].get_pos()`, `.update_agent_pos(...)`,
`.get_state()`, `.compute_reward()` -- keeps working.  The per-env adapters in map_env.py copy
position / orientation / reward between these proxies and the device around every phase.
"""
import numpy as np

BASE_ACTIONS = {0: 'MOVE_LEFT', 1: 'MOVE_RIGHT', 2: 'MOVE_UP', 3: 'MOVE_DOWN', 4: 'STAY',
                5: 'TURN_CLOCKWISE', 6: 'TURN_COUNTERCLOCKWISE'}          # agent.py:7-13
HARVEST_ACTIONS = BASE_ACTIONS.copy()                                      # agent.py:148-149
HARVEST_ACTIONS.update({7: 'FIRE'})
CLEANUP_ACTIONS = BASE_ACTIONS.copy()                                      # agent.py:186-188
CLEANUP_ACTIONS.update({7: 'FIRE', 8: 'CLEAN'})
ACTION_CODE = {v: k for k, v in CLEANUP_ACTIONS.items()}


def return_view(grid, pos, row_size, col_size):
    """(2*row_size+1) x (2*col_size+1) window of `grid` centred on pos, '0'-padded outside the map
    (utility_funcs.py:59-114; np.pad with constant 0 on a <U1 array yields the character '0')."""
    grid = np.asarray(grid)
    r0, c0 = int(pos[0]), int(pos[1])
    view = np.full((2 * col_size + 1, 2 * row_size + 1), '0', dtype=grid.dtype)
    rows = np.arange(r0 - col_size, r0 + col_size + 1)
    cols = np.arange(c0 - row_size, c0 + row_size + 1)
    rv = (rows >= 0) & (rows < grid.shape[0])
    cv = (cols >= 0) & (cols < grid.shape[1])
    view[np.ix_(rv, cv)] = grid[np.ix_(rows[rv], cols[cv])]
    return view


class Agent(object):
    """agent.py:16-145."""

    def __init__(self, agent_id, start_pos, start_orientation, grid, row_size, col_size):
        self.agent_id = agent_id
        self.pos = np.array(start_pos)
        self.orientation = start_orientation
        self.grid = grid
        self.row_size = row_size
        self.col_size = col_size
        self.reward_this_turn = 0

    @property
    def action_space(self):
        raise NotImplementedError

    @property
    def observation_space(self):
        raise NotImplementedError

    def action_map(self, action_number):
        raise NotImplementedError

    def get_state(self):
        return return_view(self.grid, self.get_pos(), self.row_size, self.col_size)

    def compute_reward(self):
        reward = self.reward_this_turn
        self.reward_this_turn = 0
        return reward

    def set_pos(self, new_pos):
        self.pos = np.array(new_pos)

    def get_pos(self):
        return self.pos

    def translate_pos_to_egocentric_coord(self, pos):
        return [self.row_size, self.col_size] + (pos - self.get_pos())

    def set_orientation(self, new_orientation):
        self.orientation = new_orientation

    def get_orientation(self):
        return self.orientation

    def get_map(self):
        return self.grid

    def return_valid_pos(self, new_pos):
        """You can't walk through walls (agent.py:105-113)."""
        new_row, new_col = new_pos
        temp_pos = np.array(new_pos).copy()
        if self.grid[new_row, new_col] == '@':
            temp_pos = self.get_pos()
        return temp_pos

    def update_agent_pos(self, new_pos):
        old_pos = self.get_pos()
        self.set_pos(self.return_valid_pos(new_pos))
        return self.get_pos(), np.array(old_pos)

    def update_agent_rot(self, new_rot):
        self.set_orientation(new_rot)

    def hit(self, char):
        raise NotImplementedError

    def consume(self, char):
        raise NotImplementedError


class _SSDAgent(Agent):
    ACTIONS = BASE_ACTIONS

    def __init__(self, agent_id, start_pos, start_orientation, grid, view_len):
        self.view_len = view_len
        super().__init__(agent_id, start_pos, start_orientation, grid, view_len, view_len)
        self.update_agent_pos(start_pos)
        self.update_agent_rot(start_orientation)

    def action_map(self, action_number):
        return self.ACTIONS[action_number]  # KeyError for an unknown action, as in the reference

    def hit(self, char):
        if char == 'F':
            self.reward_this_turn -= 50

    def fire_beam(self, char):
        if char == 'F':
            self.reward_this_turn -= 1

    def get_done(self):
        return False

    def consume(self, char):
        if char == 'A':
            self.reward_this_turn += 1
            return ' '
        return char


class HarvestAgent(_SSDAgent):
    """agent.py:152-183."""
    ACTIONS = HARVEST_ACTIONS


class CleanupAgent(_SSDAgent):
    """agent.py:191-222."""
    ACTIONS = CLEANUP_ACTIONS
