"""BatchedSSDEnv: B independent Harvest/Cleanup environments resident in the HBM of one GPU.

Tensor-level host API over the C-ABI (include/ssd_b200.h).  PyTorch only provides device memory
and the current stream; every transition and every observation is produced by the sm_100a
kernels of libssd_b200.so.  The RLlib-shaped dict API of the reference lives in
sequential_social_dilemma_games_b200.envs (HarvestEnv / CleanupEnv / MapEnv).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .config import make_config  # noqa: F401  (re-exported; lives in config.py so that CPU-only tools never load the CUDA library)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class BatchedSSDEnv(object):
    """num_envs copies of one game stepped in lock-step on `device`.

    actions  int8  [B, N]  (-1 = agent absent from the action dict)
    obs      uint8 [B, N, V, V, 3]   (the reference's float64 obs is (obs - 128.0) / 255.0)
    rewards  int32 [B, N]
    """

    def __init__(self, cfg, num_envs, device="cuda:0", seed=0, env_id_offset=0, envs_per_cta=0):
        if isinstance(cfg, str):
            cfg = make_config(cfg)
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedSSDEnv needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.cfg = cfg
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("device must be a CUDA device")
        self.device_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.env_id_offset = int(env_id_offset)
        self._keep = dict(base_map=np.ascontiguousarray(cfg.base_map), lut=np.ascontiguousarray(cfg.colour_lut),
                          hp=np.ascontiguousarray(cfg.harvest_spawn_prob), ap=np.ascontiguousarray(cfg.cleanup_apple_prob),
                          wp=np.ascontiguousarray(cfg.cleanup_waste_prob), sp=np.ascontiguousarray(cfg.spawn_points))
        k = self._keep
        c = _lib.SsdConfig(abi_version=_lib.ABI_VERSION, kind=cfg.kind, height=cfg.height, width=cfg.width,
                           num_agents=cfg.num_agents, view_radius=cfg.view_size, beam_len=cfg.beam_length,
                           num_envs=self.num_envs, device=self.device_index, envs_per_cta=envs_per_cta,
                           env_id_offset=self.env_id_offset,
                           base_map=k["base_map"].ctypes.data, color_lut=k["lut"].ctypes.data,
                           harvest_spawn_prob=k["hp"].ctypes.data, cleanup_apple_prob=k["ap"].ctypes.data,
                           cleanup_waste_prob=k["wp"].ctypes.data, potential_waste_area=cfg.potential_waste_area,
                           num_spawn_points=len(k["sp"]), spawn_points=k["sp"].ctypes.data if len(k["sp"]) else None)
        h = C.c_void_p()
        _lib.check(_lib.lib.ssd_create(C.byref(c), C.byref(h)))
        self._h = h
        self.seed(seed)
        B, N = self.num_envs, cfg.num_agents
        self.obs_shape = (B,) + cfg.obs_shape
        self._obs = None
        self._rew = None

    # ------------------------------------------------------------------ plumbing
    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            _lib.lib.ssd_destroy(h)
            self._h = None

    close = __del__

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _obs_buf(self, out):
        if out is not None:
            assert out.dtype == torch.uint8 and out.is_contiguous() and tuple(out.shape) == self.obs_shape and out.device == self._dev()
            return out
        if self._obs is None:
            self._obs = torch.empty(self.obs_shape, dtype=torch.uint8, device=self.device)
        return self._obs

    def _dev(self):
        return torch.device("cuda", self.device_index)

    def _as(self, x, dtype, shape):
        t = torch.as_tensor(x, dtype=dtype, device=self.device).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError("expected shape %s, got %s" % (tuple(shape), tuple(t.shape)))
        return t

    # ------------------------------------------------------------------ RNG
    def seed(self, seed, t=0):
        """Philox key and step counter (np.random.seed / random.seed analogue)."""
        self._seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        _lib.check(_lib.lib.ssd_seed(self._h, self._seed, int(t)))

    @property
    def t(self):
        v = C.c_uint32()
        _lib.check(_lib.lib.ssd_get_counter(self._h, C.byref(v)))
        return v.value

    # ------------------------------------------------------------------ MapEnv.reset / step
    def reset(self, mask=None, out=None, render=True):
        """MapEnv.reset (map_env.py:214-249) for all envs, or those with mask != 0."""
        m = None if mask is None else self._as(mask, torch.uint8, (self.num_envs,))
        obs = self._obs_buf(out) if render else None
        _lib.check(_lib.lib.ssd_reset(self._h, _ptr(m), _ptr(obs), self._stream()))
        return obs

    def reset_rows(self, rows, out=None, render=True):
        """MapEnv.reset of the listed envs only (ssd_reset_rows): one launch over len(rows) warps.  `rows` is a
        sequence of ints or an int32 tensor; returns the full observation tensor (only those rows are rewritten)."""
        obs = self._obs_buf(out) if render else None
        if torch.is_tensor(rows):
            r = rows.to(device=self.device, dtype=torch.int32).contiguous()
            _lib.check(_lib.lib.ssd_reset_rows(self._h, _ptr(r), int(r.numel()), _ptr(obs), self._stream()))
        else:
            r = np.ascontiguousarray(rows, dtype=np.int32).reshape(-1)
            _lib.check(_lib.lib.ssd_reset_rows(self._h, r.ctypes.data, int(r.size), _ptr(obs), self._stream()))
            torch.cuda.current_stream(self.device).synchronize()  # the host list was copied asynchronously
        return obs

    def step(self, actions, action_order=None, tape=None, out=None, reward_out=None, render=True, phases=None):
        """MapEnv.step (map_env.py:152-212).  Returns (obs, rewards); dones are always False in the
        reference (agent.py:174,209) and infos empty."""
        B, N = self.num_envs, self.cfg.num_agents
        a = self._as(actions, torch.int8, (B, N))
        o = None if action_order is None else self._as(action_order, torch.uint8, (B, N))
        obs = self._obs_buf(out) if render else None
        if reward_out is None:
            if self._rew is None:
                self._rew = torch.empty((B, N), dtype=torch.int32, device=self.device)
            reward_out = self._rew
        tp, keep = None, None
        if tape is not None:
            mo = self._as(tape["move_order"], torch.uint8, (B, N))
            u = torch.as_tensor(tape["uniforms"], dtype=torch.float64, device=self.device).contiguous()
            assert u.dim() == 2 and u.shape[0] == B
            wo = tape.get("waste_order")
            if wo is not None:
                if not torch.is_tensor(wo):  # uint16 cell ids travel as their int16 bit pattern
                    wo = torch.from_numpy(np.ascontiguousarray(wo, dtype=np.uint16).view(np.int16))
                wo = wo.to(self.device).contiguous()
                assert wo.element_size() == 2 and tuple(wo.shape) == (B, len(self.cfg.waste_points))
            nd = tape.get("n_draws_out")
            if nd is not None:
                assert nd.dtype == torch.int32 and nd.is_contiguous() and tuple(nd.shape) == (B,) and nd.is_cuda
            keep = (mo, u, wo, nd)
            tp = _lib.SsdTape(move_order=mo.data_ptr(), uniforms=u.data_ptr(), u_stride=u.shape[1],
                              waste_order=wo.data_ptr() if wo is not None else None,
                              n_draws_out=nd.data_ptr() if nd is not None else None)
        tpp = C.byref(tp) if tp is not None else None
        if phases is None:
            _lib.check(_lib.lib.ssd_step(self._h, _ptr(a), _ptr(o), tpp, _ptr(obs), _ptr(reward_out), self._stream()))
        else:
            _lib.check(_lib.lib.ssd_step_phases(self._h, int(phases), _ptr(a), _ptr(o), tpp, _ptr(obs),
                                                _ptr(reward_out), self._stream()))
        del keep
        return obs, reward_out

    def rollout(self, actions, obs_ring=None, reward_out=None):
        """T steps with actions that all exist up front (int8 [T, B, N] on the device): ssd_rollout.  Observations of
        step s land in obs_ring[s % R] (uint8 [R, B, N, V, V, 3], default R = 1), rewards in reward_out [T, B, N]."""
        B, N = self.num_envs, self.cfg.num_agents
        assert actions.is_cuda and actions.dtype == torch.int8 and actions.is_contiguous() and tuple(actions.shape[1:]) == (B, N)
        T = int(actions.shape[0])
        if obs_ring is None:
            obs_ring = torch.empty((1,) + tuple(self.obs_shape), dtype=torch.uint8, device=self.device)
        assert obs_ring.is_cuda and obs_ring.dtype == torch.uint8 and obs_ring.is_contiguous() and tuple(obs_ring.shape[1:]) == tuple(self.obs_shape)
        if reward_out is None:
            reward_out = torch.empty((T, B, N), dtype=torch.int32, device=self.device)
        assert reward_out.is_cuda and reward_out.dtype == torch.int32 and reward_out.is_contiguous() and tuple(reward_out.shape) == (T, B, N)
        _lib.check(_lib.lib.ssd_rollout(self._h, T, _ptr(actions), _ptr(obs_ring), int(obs_ring.shape[0]), _ptr(reward_out), self._stream()))
        return obs_ring, reward_out

    def render(self, rotate=True, out=None):
        obs = self._obs_buf(out)
        _lib.check(_lib.lib.ssd_render(self._h, int(bool(rotate)), _ptr(obs), self._stream()))
        return obs

    def render_map(self, out=None):
        """uint8 [B, H, W, 3] frames of the whole map with the agents painted (map_to_colors of
        get_map_with_agents, map_env.py:280-339): what rollout.py / visuallizer_rllib.py turn into videos."""
        shape = (self.num_envs, self.cfg.height, self.cfg.width, 3)
        if out is None:
            out = torch.empty(shape, dtype=torch.uint8, device=self.device)
        assert out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and tuple(out.shape) == shape
        _lib.check(_lib.lib.ssd_render_map(self._h, _ptr(out), self._stream()))
        return out

    def step_host(self, actions_host, obs_host=None, reward_host=None):
        """End-to-end step with host (ideally pinned) numpy buffers: H2D actions, fused step,
        D2H observations + rewards, all inside libssd_b200 (ssd_step_host)."""
        B, N = self.num_envs, self.cfg.num_agents
        a = np.ascontiguousarray(actions_host, dtype=np.int8)
        assert a.shape == (B, N)
        if reward_host is None:
            reward_host = np.empty((B, N), dtype=np.int32)
        assert reward_host.dtype == np.int32 and reward_host.flags.c_contiguous and reward_host.shape == (B, N)
        if obs_host is not None:
            assert obs_host.dtype == np.uint8 and obs_host.flags.c_contiguous and obs_host.shape == self.obs_shape
        _lib.check(_lib.lib.ssd_step_host(self._h, a.ctypes.data, obs_host.ctypes.data if obs_host is not None else None,
                                          reward_host.ctypes.data, self._stream()))
        return obs_host, reward_host

    # ------------------------------------------------------------------ state I/O
    def get_state(self):
        """(grid u8[B,H,W] ASCII, pos i16[B,N,2], ori u8[B,N]) as CUDA tensors."""
        B, N, c = self.num_envs, self.cfg.num_agents, self.cfg
        grid = torch.empty((B, c.height, c.width), dtype=torch.uint8, device=self.device)
        pos = torch.empty((B, N, 2), dtype=torch.int16, device=self.device)
        ori = torch.empty((B, N), dtype=torch.uint8, device=self.device)
        _lib.check(_lib.lib.ssd_get_state(self._h, _ptr(grid), _ptr(pos), _ptr(ori), self._stream()))
        return grid, pos, ori

    def set_state(self, grid, pos, ori):
        B, N, c = self.num_envs, self.cfg.num_agents, self.cfg
        g = self._as(grid, torch.uint8, (B, c.height, c.width))
        p = self._as(pos, torch.int16, (B, N, 2))
        o = self._as(ori, torch.uint8, (B, N))
        _lib.check(_lib.lib.ssd_set_state(self._h, _ptr(g), _ptr(p), _ptr(o), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()  # g/p/o may be temporaries

    def get_beams(self):
        """uint8 [B, 64]: ray lengths and beam characters of the last phase-split BEAMS call (ssd_get_beams)."""
        out = np.zeros((self.num_envs, 64), dtype=np.uint8)
        _lib.check(_lib.lib.ssd_get_beams(self._h, out.ctypes.data, self._stream()))
        return out

    def chain_steps(self, on=True):
        """SSD_OPT_CHAIN_STEPS (include/ssd_b200.h): overlap the kernels of consecutive step() calls with
        programmatic dependent launch.  Only for rollouts whose actions exist before the previous step was
        enqueued (pre-generated / scripted actions); results are unchanged."""
        _lib.check(_lib.lib.ssd_set_option(self._h, _lib.OPT_CHAIN_STEPS, int(bool(on))))
        return self

    def general_kernel_only(self, on=True):
        """SSD_OPT_GENERAL_KERNEL: step with the general kernel even where the specialised one applies (tests)."""
        _lib.check(_lib.lib.ssd_set_option(self._h, _lib.OPT_GENERAL_KERNEL, int(bool(on))))
        return self

    def stats(self):
        out = np.zeros(_lib.NUM_STATS, dtype=np.int64)
        _lib.check(_lib.lib.ssd_stats(self._h, out.ctypes.data, self._stream()))
        return dict(zip(_lib.STAT_NAMES, (int(x) for x in out)))

    @property
    def launch_count(self):
        return int(_lib.lib.ssd_launch_count(self._h))

    @property
    def envs_per_cta(self):
        return int(_lib.lib.ssd_envs_per_cta(self._h))

    @property
    def algorithmic_bytes_per_env_step(self):
        return int(_lib.lib.ssd_algorithmic_bytes_per_env_step(self._h))


def philox_selftest(ctr, key, device=0):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    _lib.check(_lib.lib.ssd_philox_selftest(device, c.ctypes.data, k.ctypes.data, out.ctypes.data))
    return out
