"""RLlib-shaped vector environment over ONE batched device state.

RLlib's sampler drives a list of sub-environments through `vector_reset` / `reset_at` / `vector_step`
(ray.rllib.env.VectorEnv; the reference only ever hands it single MultiAgentEnvs made by the env creators of
train_baseline.py:71-81 / train_moa.py:68-78).  Here the sub-environments are rows of one
`BatchedSSDEnv`: a `vector_step` is one fused kernel launch for all of them, the per-env dicts
(`{'agent-i': obs}`, rewards, dones with '__all__', infos) are built on the host from the two result
tensors.  Randomness comes from the production Philox streams (seed, global env id, step), not from the
module-level numpy / random generators the per-env adapters replay.

Semantics kept from MapEnv.step / reset (map_env.py:152-249): agents missing from an action dict do not
act; the iteration order of the dict is the action order (it decides move priorities and who fires first);
observations are float64 `(rgb - 128) / 255` (or the uint8 array when `uint8_obs=True`); `reset` returns
un-rotated views; with `return_agent_actions=True` every observation is the dict
{curr_obs, other_agent_actions (string-sorted ids), visible_agents (all ones, map_env.py:749-769)}.
`horizon` reproduces RLlib's episode horizon (train_baseline.py:131): dones['__all__'] turns True after
that many steps and the caller resets the row with `reset_at`.
"""
import numpy as np
import torch

from ..batched import BatchedSSDEnv, make_config
from .harvest import HarvestEnv
from .cleanup import CleanupEnv
from .spaces import Discrete

_OBS_LUT = (np.arange(256) - 128.0) / 255.0   # map_env.py:199 on every possible uint8


class SSDVectorEnv(object):
    def __init__(self, game, num_envs, num_agents=5, ascii_map=None, view_size=7, device="cuda:0", seed=0,
                 return_agent_actions=False, horizon=None, uint8_obs=False, env_id_offset=0):
        self.cfg = make_config(game, num_agents=num_agents, view_size=view_size, ascii_map=ascii_map)
        self.num_envs, self.num_agents = int(num_envs), int(num_agents)
        self.engine = BatchedSSDEnv(self.cfg, self.num_envs, device=device, seed=seed, env_id_offset=env_id_offset)
        self.agent_ids = ['agent-%d' % i for i in range(self.num_agents)]
        self._index = {a: i for i, a in enumerate(self.agent_ids)}
        self._sorted_ids = sorted(self.agent_ids)
        self.return_agent_actions = bool(return_agent_actions)
        self.horizon = horizon
        self.uint8_obs = bool(uint8_obs)
        self._t = np.zeros(self.num_envs, dtype=np.int64)
        proto = (HarvestEnv if game.lower() == "harvest" else CleanupEnv)
        self.action_space = Discrete(proto.NUM_ACTIONS)
        self.view_len = view_size
        self.observation_space = proto.observation_space.fget(self)
        self._act = torch.empty((self.num_envs, self.num_agents), dtype=torch.int8, device=self.engine.device)
        self._order = torch.empty((self.num_envs, self.num_agents), dtype=torch.uint8, device=self.engine.device)

    # ------------------------------------------------------------------ helpers
    def _wrap(self, obs_u8, b, actions=None):
        out = {}
        for i, aid in enumerate(self.agent_ids):
            o = obs_u8[b, i] if self.uint8_obs else _OBS_LUT[obs_u8[b, i]]
            if self.return_agent_actions:
                if actions is None:
                    prev = np.zeros(self.num_agents - 1, dtype=np.int64)
                else:
                    prev = np.array([actions[k] for k in sorted(actions.keys()) if k != aid]).astype(np.int64)
                o = {"curr_obs": o, "other_agent_actions": prev, "visible_agents": np.ones(self.num_agents - 1, dtype=np.int64)}
            out[aid] = o
        return out

    # ------------------------------------------------------------------ VectorEnv surface
    def vector_reset(self):
        obs = self.engine.reset().cpu().numpy()
        self._t[:] = 0
        return [self._wrap(obs, b) for b in range(self.num_envs)]

    def reset_at(self, index):
        obs = self.engine.reset_rows([int(index)])[index:index + 1].cpu().numpy()  # one warp, not a pass over all rows
        self._t[index] = 0
        return self._wrap(obs, 0)

    def vector_step(self, actions):
        """actions: list (one per sub-env) of {agent_id: int}.  Returns (obs, rewards, dones, infos) lists."""
        assert len(actions) == self.num_envs
        n = self.num_agents
        act = np.full((self.num_envs, n), -1, dtype=np.int8)
        order = np.tile(np.arange(n, dtype=np.uint8), (self.num_envs, 1))
        custom_order = False
        for b, d in enumerate(actions):
            present = []
            for aid, a in d.items():
                i = self._index[aid]          # KeyError for an unknown agent id
                if not 0 <= int(a) < self.action_space.n:
                    raise KeyError(a)         # agent.action_map raises KeyError for an unknown action (agent.py:162)
                act[b, i] = int(a)
                present.append(i)
            if present != sorted(present):
                custom_order = True
            order[b] = present + [i for i in range(n) if i not in present]
        self._act.copy_(torch.from_numpy(act))
        if custom_order:
            self._order.copy_(torch.from_numpy(order))
        obs, rew = self.engine.step(self._act, action_order=self._order if custom_order else None)
        obs, rew = obs.cpu().numpy(), rew.cpu().numpy()
        self._t += 1
        out_obs, out_rew, out_done, out_info = [], [], [], []
        for b in range(self.num_envs):
            out_obs.append(self._wrap(obs, b, actions[b]))
            out_rew.append({aid: int(rew[b, i]) for i, aid in enumerate(self.agent_ids)})
            dones = {aid: False for aid in self.agent_ids}   # agent.py:174,209: never done
            dones["__all__"] = bool(self.horizon is not None and self._t[b] >= self.horizon)
            out_done.append(dones)
            out_info.append({})
        return out_obs, out_rew, out_done, out_info

    def get_unwrapped(self):
        return []

    # ------------------------------------------------------------------ tensors for an on-GPU policy
    def step_tensors(self, actions, out=None, reward_out=None):
        """int8 [B, N] actions on the device -> (uint8 obs [B, N, V, V, 3], int32 rewards [B, N]) on the device."""
        self._t += 1
        return self.engine.step(actions, out=out, reward_out=reward_out)

    def close(self):
        self.engine.close()
