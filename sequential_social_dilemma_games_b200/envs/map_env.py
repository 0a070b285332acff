"""MapEnv with the reference's per-environment dict API, backed by the CUDA step kernels.

Drop-in surface of social_dilemmas/envs/map_env.py: `reset()` / `step({agent_id: int})` returning
per-agent obs / reward / done / info dicts, `world_map`, `agents`, `test_map`, `beam_pos`,
`get_map_with_agents()`, `map_to_colors()`, `update_map()`, the `custom_reset` /
`custom_action` / `custom_map_update` / `setup_agents` hooks, agent ids and action enums.

How a step runs: the host objects (world_map, agent proxies) are uploaded with ssd_set_state, the
phases of MapEnv.step run on the GPU through ssd_step_phases, and the results are downloaded again
-- one environment per object, so this layer is about fidelity, not speed (the throughput API is
batched.BatchedSSDEnv).  The random numbers are drawn exactly where the reference draws them,
from `np.random` / `random` (np.random.shuffle for the move priority map_env.py:422,
np.random.rand per eligible spawn cell harvest.py:101 / cleanup.py:139,150, random.shuffle of the
waste points cleanup.py:145, random.shuffle + np.random.randint in reset map_env.py:656,666) and
handed to the kernels as a replay tape, so that a program seeded like a reference program sees
the same trajectory, bit for bit (tests/test_adapter_gpu.py replays the golden fixtures this way).
A subclass that overrides a hook gets its Python code called at the reference's call site, with
the remaining phases still executed on the device.
"""
import random
from collections import defaultdict

import numpy as np

from .. import _lib
from ..config import DEFAULT_COLOURS as _DEFAULT_COLOURS
from ..config import EnvConfig, KIND_CLEANUP, KIND_HARVEST, KIND_PLAIN
from .agent import ACTION_CODE

try:  # the reference derives from RLlib's MultiAgentEnv (map_env.py:9,60); keep that when ray is around
    from ray.rllib.env import MultiAgentEnv  # pragma: no cover
except Exception:  # pragma: no cover
    class MultiAgentEnv(object):
        pass

ACTIONS = {'MOVE_LEFT': [-1, 0], 'MOVE_RIGHT': [1, 0], 'MOVE_UP': [0, -1], 'MOVE_DOWN': [0, 1], 'STAY': [0, 0],
           'TURN_CLOCKWISE': [[0, -1], [1, 0]], 'TURN_COUNTERCLOCKWISE': [[0, 1], [-1, 0]],
           'FIRE': 5, 'CLEAN': 5}                                             # map_env.py:11-17, harvest.py:11, cleanup.py:11-12
ORIENTATIONS = {'LEFT': [-1, 0], 'RIGHT': [1, 0], 'UP': [0, -1], 'DOWN': [0, 1]}  # map_env.py:19-22
# one mutable module-level table shared by every env without an explicit color_map (map_env.py:92)
DEFAULT_COLOURS = {k: list(v) for k, v in _DEFAULT_COLOURS.items()}
DEFAULT_COLOURS[''] = [180, 180, 180]

ORI_CODE = {'UP': 0, 'RIGHT': 1, 'DOWN': 2, 'LEFT': 3}
ORI_NAME = {v: k for k, v in ORI_CODE.items()}
_OBS_LUT = (np.arange(256) - 128.0) / 255.0   # map_env.py:199 on every possible uint8


class MapEnv(MultiAgentEnv):
    KIND = KIND_PLAIN
    VIEW_SIZE = 2      # subclasses: HARVEST_VIEW_SIZE / CLEANUP_VIEW_SIZE

    def __init__(self, ascii_map, num_agents=1, render=True, color_map=None, return_agent_actions=False,
                 device="cuda:0"):
        self.num_agents = num_agents
        self.base_map = self.ascii_to_numpy(ascii_map)
        self._ascii_map = [''.join(str(ch) for ch in row) for row in ascii_map]  # list of strings or a 2-D character array (tests/test_envs.py:36-44)
        self.return_agent_actions = return_agent_actions
        if self.return_agent_actions:
            self.prev_actions = defaultdict(lambda: [0] * self.num_agents)
        self.world_map = np.full((len(self.base_map), len(self.base_map[0])), ' ')
        self.beam_pos = []
        self.agents = {}
        self.pos_dict = {}
        self.color_map = color_map if color_map is not None else DEFAULT_COLOURS
        self.spawn_points = []
        self.wall_points = []
        for row in range(self.base_map.shape[0]):
            for col in range(self.base_map.shape[1]):
                if self.base_map[row, col] == 'P':
                    self.spawn_points.append([row, col])
                elif self.base_map[row, col] == '@':
                    self.wall_points.append([row, col])
        self._device = device
        self._engines = {}
        self.setup_agents()

    # ------------------------------------------------------------------ hooks (map_env.py:104-129)
    def custom_reset(self):
        pass

    def custom_action(self, agent, action):
        pass

    def custom_map_update(self):
        pass

    def setup_agents(self):
        raise NotImplementedError

    def ascii_to_numpy(self, ascii_list):
        arr = np.full((len(ascii_list), len(ascii_list[0])), ' ')
        for row in range(arr.shape[0]):
            for col in range(arr.shape[1]):
                arr[row, col] = ascii_list[row][col]
        return arr

    # ------------------------------------------------------------------ device plumbing
    def _view_size(self):
        for a in self.agents.values():
            return int(a.row_size)
        return int(self.VIEW_SIZE)

    def _engine(self, view=None):
        """BatchedSSDEnv of one env for the current number of agents and a view size (a ghost agent parked on
        a wall cell stands in when the env has none: it never acts, blocks nothing and is never observed)."""
        from ..batched import BatchedSSDEnv
        n = max(1, len(self.agents))
        key = (n, self._view_size() if view is None else int(view))
        if key not in self._engines:
            cfg = EnvConfig(self.KIND, self._ascii_map, n, view_size=key[1],
                            colour_map={k: v for k, v in self.color_map.items() if len(k) == 1},
                            **self._config_kwargs())
            self._engines[key] = BatchedSSDEnv(cfg, 1, device=self._device)
        return self._engines[key]

    def _config_kwargs(self):
        return {}

    def _upload(self, eng):
        agents = list(self.agents.values())
        n = eng.cfg.num_agents
        pos = np.zeros((1, n, 2), dtype=np.int16)
        ori = np.zeros((1, n), dtype=np.uint8)
        for i, a in enumerate(agents):
            pos[0, i] = a.pos
            ori[0, i] = ORI_CODE[a.orientation]
        grid = np.vectorize(ord)(self.world_map).astype(np.uint8)[None]
        eng.set_state(grid, pos, ori)

    def _download(self, eng, rewards=None):
        g, p, o = eng.get_state()
        g, p, o = g.cpu().numpy()[0], p.cpu().numpy()[0], o.cpu().numpy()[0]
        self.world_map = g.view('S1').astype('<U1').reshape(g.shape)
        for i, a in enumerate(self.agents.values()):
            a.set_pos(p[i].astype(np.int64))
            a.set_orientation(ORI_NAME[int(o[i])])
            if rewards is not None:
                a.reward_this_turn += int(rewards[i])

    def _run_phases(self, phases, actions=None, order=None, move_order=None, uniforms=None, waste_order=None,
                    render=False, rotate=True, view=None):
        """Upload host state, run `phases` on the device with a replay tape, download the results."""
        import torch
        eng = self._engine(view)
        n = eng.cfg.num_agents
        self._upload(eng)
        act = np.full((1, n), -1, dtype=np.int8) if actions is None else actions
        mo = np.full((1, n), 255, dtype=np.uint8) if move_order is None else move_order
        u = np.zeros((1, 1), dtype=np.float64) if uniforms is None else uniforms
        nd = torch.zeros(1, dtype=torch.int32, device=eng.device)
        tape = dict(move_order=mo, uniforms=u, n_draws_out=nd)
        nw = len(eng.cfg.waste_points)
        if nw:
            tape["waste_order"] = waste_order if waste_order is not None else np.array(
                [[int(r) * eng.cfg.width + int(c) for r, c in eng.cfg.waste_points]], dtype=np.uint16)
        rew = torch.zeros((1, n), dtype=torch.int32, device=eng.device)
        if phases == _lib.PHASE_RENDER and not rotate:
            obs = eng.render(rotate=False)
        else:
            obs, _ = eng.step(act, action_order=order, tape=tape, reward_out=rew, render=render, phases=phases)
        self._download(eng, rew.cpu().numpy()[0])
        return (obs.cpu().numpy()[0] if obs is not None else None), int(nd.item())

    # ------------------------------------------------------------------ MapEnv.step (map_env.py:152-212)
    def step(self, actions):
        self.beam_pos = []
        agent_actions = {}
        for agent_id, action in actions.items():
            agent_actions[agent_id] = self.agents[agent_id].action_map(action)

        self.update_moves(agent_actions)     # moves + consume on the device
        self.update_custom_moves(agent_actions)
        self.custom_map_update()
        map_with_agents = self.get_map_with_agents()

        obs_u8 = self._render_obs(rotate=True)
        observations, rewards, dones, info = {}, {}, {}, {}
        for i, agent in enumerate(self.agents.values()):
            agent.grid = map_with_agents
            rgb_arr = _OBS_LUT[obs_u8[i]]
            if self.return_agent_actions:
                prev_actions = np.array([actions[key] for key in sorted(actions.keys())
                                         if key != agent.agent_id]).astype(np.int64)
                observations[agent.agent_id] = {"curr_obs": rgb_arr, "other_agent_actions": prev_actions,
                                                "visible_agents": self.find_visible_agents(agent.agent_id)}
            else:
                observations[agent.agent_id] = rgb_arr
            rewards[agent.agent_id] = agent.compute_reward()
            dones[agent.agent_id] = agent.get_done()
        dones["__all__"] = np.any(list(dones.values()))
        return observations, rewards, dones, info

    def reset(self):
        """map_env.py:214-249."""
        self.beam_pos = []
        self.agents = {}
        self.setup_agents()
        self.reset_map()
        self.custom_map_update()
        map_with_agents = self.get_map_with_agents()
        obs_u8 = self._render_obs(rotate=False)   # reset() does not rotate the view (map_env.py:239-240)
        observations = {}
        for i, agent in enumerate(self.agents.values()):
            agent.grid = map_with_agents
            rgb_arr = _OBS_LUT[obs_u8[i]]
            if self.return_agent_actions:
                prev_actions = np.array([0 for _ in range(self.num_agents - 1)]).astype(np.int64)
                observations[agent.agent_id] = {"curr_obs": rgb_arr, "other_agent_actions": prev_actions,
                                                "visible_agents": self.find_visible_agents(agent.agent_id)}
            else:
                observations[agent.agent_id] = rgb_arr
        return observations

    def _render_obs(self, rotate):
        """uint8 observations of all agents, as a list (agents may have different view sizes: the reference renders
        every agent with its own row_size, agent.py:76-78)."""
        if not self.agents:
            return []
        agents = list(self.agents.values())
        out = [None] * len(agents)
        for view in sorted({int(a.row_size) for a in agents}):
            if self._beams_from_device and view == self._view_size():  # the engine that ran the beam phase holds the beams
                obs, _ = self._run_phases(_lib.PHASE_RENDER, render=True, rotate=rotate, view=view,
                                          order=self._step_order if rotate else None)
            else:  # beams from an overridden custom_action, or another view size: render without, paint beam_pos on the host
                eng = self._engine(view)
                self._upload(eng)
                obs = eng.render(rotate=rotate).cpu().numpy()[0]
                self._paint_beams_host(obs, rotate, view)
            for i, a in enumerate(agents):
                if int(a.row_size) == view:
                    out[i] = obs[i]
        return out

    _beams_from_device = True
    _step_act = None
    _step_order = None

    # ------------------------------------------------------------------ update_moves (map_env.py:357-543)
    def _action_arrays(self, agent_actions):
        ids = list(self.agents.keys())
        n = max(1, len(ids))
        act = np.full((1, n), -1, dtype=np.int8)
        present = []
        for agent_id, name in agent_actions.items():
            i = ids.index(agent_id)
            act[0, i] = ACTION_CODE[name]
            present.append(i)
        order = np.array([present + [i for i in range(n) if i not in present]], dtype=np.uint8)
        return act, order, present

    def update_moves(self, agent_actions):
        """Moves, rotations and conflict resolution, followed by the consume loop (map_env.py:176-181).
        np.random.shuffle is called exactly as the reference calls it: on the list of movers in
        action-dict order, only when there is at least one (map_env.py:415-423)."""
        act, order, present = self._action_arrays(agent_actions)
        movers = [i for i in present if 0 <= act[0, i] <= 4]
        mo = np.full((1, act.shape[1]), 255, dtype=np.uint8)
        if movers:
            np.random.shuffle(movers)
            mo[0, :len(movers)] = movers
        self._step_act, self._step_order = act, order
        from .agent import _SSDAgent
        stock = all(type(a).consume is _SSDAgent.consume for a in self.agents.values())
        self._run_phases(_lib.PHASE_MOVES | (_lib.PHASE_CONSUME if stock else 0), actions=act, order=order, move_order=mo)
        if not stock:  # agents with their own consume(): the reference's loop, map_env.py:178-181
            for agent in self.agents.values():
                pos = agent.get_pos()
                self.world_map[pos[0], pos[1]] = agent.consume(self.world_map[pos[0], pos[1]])

    def _custom_action_overridden(self):
        return type(self).custom_action is not getattr(type(self), "_device_custom_action", None)

    def update_custom_moves(self, agent_actions):
        """map_env.py:545-552.  Default hook -> the device beam phase; an overridden custom_action is
        called per firing agent in action-dict order and its updates applied immediately."""
        if not self._custom_action_overridden():
            self._beams_from_device = True
            if self.KIND == KIND_PLAIN or not self.agents:
                return
            self._run_phases(_lib.PHASE_BEAMS, actions=self._step_act, order=self._step_order)
            self._collect_beams()
            return
        self._beams_from_device = False
        for agent_id, action in agent_actions.items():
            if 'MOVE' not in action and 'STAY' not in action and 'TURN' not in action:
                updates = self.custom_action(self.agents[agent_id], action)
                if len(updates) > 0:
                    self.update_map(updates)

    def _collect_beams(self):
        """Rebuild beam_pos (map_env.py:648) from the ray lengths the device recorded."""
        eng = self._engine()
        rec = eng.get_beams()[0]
        agents = list(self.agents.values())
        for k in range(len(agents)):
            ch = rec[48 + k]
            if not ch:
                continue
            agent = agents[int(self._step_order[0, k])]
            d = np.array(ORIENTATIONS[agent.orientation])
            right = np.array([-d[1], d[0]])
            starts = [agent.pos, agent.pos + right - d, agent.pos - right - d]
            for s in range(3):
                cell = np.array(starts[s]) + d
                for _ in range(int(rec[k * 3 + s])):
                    self.beam_pos.append((int(cell[0]), int(cell[1]), chr(ch)))
                    cell = cell + d

    def _paint_beams_host(self, obs, rotate, view):
        V = 2 * view + 1
        for i, agent in enumerate(self.agents.values()):
            k = {'UP': 0, 'LEFT': 1, 'DOWN': 2, 'RIGHT': 3}[agent.orientation] if rotate else 0
            for (r, c, ch) in self.beam_pos:
                vi, vj = r - agent.pos[0] + view, c - agent.pos[1] + view
                if not (0 <= vi < V and 0 <= vj < V):
                    continue
                if k == 1:
                    vi, vj = V - 1 - vj, vi
                elif k == 2:
                    vi, vj = V - 1 - vi, V - 1 - vj
                elif k == 3:
                    vi, vj = vj, V - 1 - vi
                obs[i, vi, vj] = self.color_map[ch]

    # ------------------------------------------------------------------ spawning helper for subclasses
    def _device_spawn(self, waste_order=None):
        """Run the device spawn pass on the current host state with np.random.rand values as the
        tape, advance np.random by exactly the number of draws the reference would have made, and
        return the new cells as [(row, col, char)] without applying them (harvest.py:75-104,
        cleanup.py:132-154 return such lists)."""
        eng = self._engine()
        before = self.world_map.copy()
        state = np.random.get_state()
        u = np.random.rand(1, max(1, eng.cfg.max_draws))
        np.random.set_state(state)
        _, n_draws = self._run_phases(_lib.PHASE_SPAWN, uniforms=u, waste_order=waste_order)
        if n_draws:
            np.random.rand(n_draws)  # same stream position as n_draws calls of np.random.rand(1)
        after = self.world_map
        self.world_map = before
        rr, cc = np.nonzero(after != before)
        return [(int(r), int(c), str(after[r, c])) for r, c in zip(rr, cc)]

    # ------------------------------------------------------------------ host-side surface of the reference
    @property
    def agent_pos(self):
        return [agent.get_pos().tolist() for agent in self.agents.values()]

    @property
    def test_map(self):
        """map_env.py:257-278."""
        grid = np.copy(self.world_map)
        for agent in self.agents.values():
            if 0 <= agent.pos[0] < grid.shape[0] and 0 <= agent.pos[1] < grid.shape[1]:
                grid[agent.pos[0], agent.pos[1]] = 'P'
        for beam_pos in self.beam_pos:
            grid[beam_pos[0], beam_pos[1]] = beam_pos[2]
        return grid

    def get_map_with_agents(self):
        """map_env.py:280-302 (the agent character is the last digit of the id plus one, cut to one
        character by the <U1 array: agent-9 shows as '1')."""
        grid = np.copy(self.world_map)
        for agent_id, agent in self.agents.items():
            char_id = str(int(agent_id[-1]) + 1)
            if 0 <= agent.pos[0] < grid.shape[0] and 0 <= agent.pos[1] < grid.shape[1]:
                grid[agent.pos[0], agent.pos[1]] = char_id
        for beam_pos in self.beam_pos:
            grid[beam_pos[0], beam_pos[1]] = beam_pos[2]
        return grid

    def check_agent_map(self, agent_map):
        unique, counts = np.unique(agent_map, return_counts=True)
        count_dict = dict(zip(unique, counts))
        for i in range(self.num_agents):
            if count_dict[str(i + 1)] != 1:
                print('Error! Wrong number of agent', i, 'in map!')
                return False
        return True

    def map_to_colors(self, map=None, color_map=None):
        """map_env.py:316-339."""
        if map is None:
            map = self.get_map_with_agents()
        if color_map is None:
            color_map = self.color_map
        rgb_arr = np.zeros((map.shape[0], map.shape[1], 3), dtype=int)
        for ch in np.unique(map):
            rgb_arr[map == ch] = color_map[str(ch)]
        return rgb_arr

    def render(self, filename=None):
        import matplotlib.pyplot as plt  # only needed here, as in the reference (map_env.py:341-355)
        rgb_arr = self.map_to_colors(self.get_map_with_agents())
        plt.imshow(rgb_arr, interpolation='nearest')
        if filename is None:
            plt.show()
        else:
            plt.savefig(filename)

    def update_map(self, new_points):
        for i in range(len(new_points)):
            row, col, char = new_points[i]
            self.world_map[row, col] = char

    def reset_map(self):
        self.world_map = np.full((len(self.base_map), len(self.base_map[0])), ' ')
        self.build_walls()
        self.custom_reset()

    def build_walls(self):
        for row, col in self.wall_points:
            self.world_map[row, col] = '@'

    def spawn_point(self):
        """map_env.py:651-662: shuffle the persistent list, take the LAST free entry."""
        spawn_index = 0
        is_free_cell = False
        curr_agent_pos = [agent.get_pos().tolist() for agent in self.agents.values()]
        random.shuffle(self.spawn_points)
        for i, spawn_point in enumerate(self.spawn_points):
            if [spawn_point[0], spawn_point[1]] not in curr_agent_pos:
                spawn_index = i
                is_free_cell = True
        assert is_free_cell, 'There are not enough spawn points! Check your map?'
        return np.array(self.spawn_points[spawn_index])

    def spawn_rotation(self):
        rand_int = np.random.randint(len(ORIENTATIONS.keys()))
        return list(ORIENTATIONS.keys())[rand_int]

    def rotate_view(self, orientation, view):
        if orientation == 'UP':
            return view
        elif orientation == 'LEFT':
            return np.rot90(view, k=1, axes=(0, 1))
        elif orientation == 'DOWN':
            return np.rot90(view, k=2, axes=(0, 1))
        elif orientation == 'RIGHT':
            return np.rot90(view, k=3, axes=(0, 1))
        raise ValueError('Orientation {} is not valid'.format(orientation))

    def rotate_action(self, action_vec, orientation):
        if orientation == 'UP':
            return action_vec
        elif orientation == 'LEFT':
            return self.rotate_left(action_vec)
        elif orientation == 'RIGHT':
            return self.rotate_right(action_vec)
        return self.rotate_left(self.rotate_left(action_vec))

    def rotate_left(self, action_vec):
        return np.dot(ACTIONS['TURN_COUNTERCLOCKWISE'], action_vec)

    def rotate_right(self, action_vec):
        return np.dot(ACTIONS['TURN_CLOCKWISE'], action_vec)

    def update_rotation(self, action, curr_orientation):
        cw = ['UP', 'RIGHT', 'DOWN', 'LEFT']
        step = 1 if action == 'TURN_CLOCKWISE' else -1
        return cw[(cw.index(curr_orientation) + step) % 4]

    def test_if_in_bounds(self, pos):
        return 0 <= pos[0] < self.world_map.shape[0] and 0 <= pos[1] < self.world_map.shape[1]

    def find_visible_agents(self, agent_id):
        """map_env.py:749-770 -- including its quirk: every other agent is tested at the CALLER's own
        position, so the result is all ones."""
        agent_pos = self.agents[agent_id].get_pos()
        upper_lim = int(agent_pos[0] + self.agents[agent_id].row_size)
        lower_lim = int(agent_pos[0] - self.agents[agent_id].row_size)
        left_lim = int(agent_pos[1] - self.agents[agent_id].col_size)
        right_lim = int(agent_pos[1] + self.agents[agent_id].col_size)
        other_agent_pos = [self.agents[agent_id].get_pos() for other_agent_id in sorted(self.agents.keys())
                           if other_agent_id != agent_id]
        return np.array([1 if (lower_lim <= agent_tup[0] <= upper_lim and left_lim <= agent_tup[1] <= right_lim)
                         else 0 for agent_tup in other_agent_pos])


MapEnv._device_custom_action = MapEnv.custom_action
