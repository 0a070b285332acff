"""TEST INFRASTRUCTURE ONLY (see oracle/README.md): ctypes binding of oracle/libssd_oracle.so.

Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
as the checker / the timed CPU baseline.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libssd_oracle.so")
NUM_STATS = 8
STAT_NAMES = ("env_steps", "reward_sum", "apples_eaten", "fires", "hits", "cleaned",
              "apples_spawned", "waste_spawned")


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("ssd_oracle.c", "ssd_oracle.h", "Makefile")]
    if (force or not os.path.exists(_LIB_PATH)
            or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libssd_oracle.so"])
    return _LIB_PATH


class _Tape(C.Structure):
    _fields_ = [("move_order", C.c_void_p), ("uniforms", C.c_void_p), ("u_stride", C.c_int32),
                ("waste_order", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int] * 6 + [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        for f in (L.orc_num_apple_points, L.orc_num_waste_points, L.orc_max_draws):
            f.argtypes = [C.c_void_p]
            f.restype = C.c_int
        L.orc_step.restype = C.c_int
        L.orc_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_reset.restype = C.c_int
        L.orc_reset.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                C.c_uint64, C.c_uint32, C.c_void_p, C.c_int]
        L.orc_render.restype = C.c_int
        L.orc_render.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                 C.c_void_p, C.c_int]
        L.orc_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _chk(a, dtype, shape):
    assert a.dtype == dtype and a.flags.c_contiguous and tuple(a.shape) == tuple(shape), \
        (a.dtype, a.shape, dtype, shape)
    return a


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(c), _p(k), _p(out))
    return out


class OracleEnv(object):
    """B independent envs stepped on the CPU; state lives in numpy arrays owned by the caller
    or (by default) by this object: grid u8[B,H,W], pos i16[B,N,2], ori u8[B,N]."""

    def __init__(self, cfg, num_envs, seed=0, env_id_offset=0, n_threads=1):
        self.cfg, self.B, self.seed, self.env_id_offset = cfg, int(num_envs), int(seed), int(env_id_offset)
        self.n_threads = n_threads
        self.t = 0
        L = lib()
        self._h = L.orc_create(cfg.kind, cfg.height, cfg.width, cfg.num_agents, cfg.view_size,
                               cfg.beam_length, _p(np.ascontiguousarray(cfg.base_map)),
                               _p(np.ascontiguousarray(cfg.colour_lut)),
                               _p(np.ascontiguousarray(cfg.harvest_spawn_prob)),
                               _p(cfg.cleanup_apple_prob), _p(cfg.cleanup_waste_prob),
                               cfg.potential_waste_area, _p(cfg.spawn_points), len(cfg.spawn_points))
        if not self._h:
            raise ValueError("orc_create rejected the configuration")
        assert L.orc_max_draws(self._h) == cfg.max_draws
        B, N = self.B, cfg.num_agents
        self.grid = np.tile(cfg.initial_grid()[None], (B, 1, 1))
        self.pos = np.zeros((B, N, 2), dtype=np.int16)
        self.ori = np.zeros((B, N), dtype=np.uint8)
        self.stats = np.zeros(NUM_STATS, dtype=np.int64)
        self.last_n_draws = np.zeros(B, dtype=np.int32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    def set_state(self, grid, pos, ori):
        self.grid[...] = np.asarray(grid, dtype=np.uint8).reshape(self.grid.shape)
        self.pos[...] = np.asarray(pos, dtype=np.int16).reshape(self.pos.shape)
        self.ori[...] = np.asarray(ori, dtype=np.uint8).reshape(self.ori.shape)

    def _new_obs(self):
        return np.zeros((self.B,) + self.cfg.obs_shape, dtype=np.uint8)

    def step(self, actions, action_order=None, tape=None, render=True, obs_out=None):
        """tape = dict(move_order u8[B,N], uniforms f64[B,K], waste_order u16[B,nw] or None).
        obs_out: optional preallocated uint8 buffer (the timed baseline reuses one)."""
        cfg, B, N = self.cfg, self.B, self.cfg.num_agents
        actions = _chk(np.ascontiguousarray(actions, dtype=np.int8), np.int8, (B, N))
        if action_order is not None:
            action_order = _chk(np.ascontiguousarray(action_order, dtype=np.uint8), np.uint8, (B, N))
        obs = (obs_out if obs_out is not None else self._new_obs()) if render else None
        rew = np.zeros((B, N), dtype=np.int32)
        tp, keep = None, []
        if tape is not None:
            mo = _chk(np.ascontiguousarray(tape["move_order"], dtype=np.uint8), np.uint8, (B, N))
            u = np.ascontiguousarray(tape["uniforms"], dtype=np.float64)
            assert u.ndim == 2 and u.shape[0] == B
            wo = tape.get("waste_order")
            if wo is not None:
                wo = _chk(np.ascontiguousarray(wo, dtype=np.uint16), np.uint16, (B, len(cfg.waste_points)))
            keep = [mo, u, wo]
            tp = _Tape(_p(mo), _p(u), u.shape[1], _p(wo))
        rc = lib().orc_step(self._h, B, _p(self.grid), _p(self.pos), _p(self.ori), _p(actions),
                            _p(action_order), C.byref(tp) if tp is not None else None, self.seed,
                            self.env_id_offset, self.t, _p(obs), _p(rew), _p(self.last_n_draws),
                            _p(self.stats), self.n_threads)
        del keep
        if rc:
            raise RuntimeError("oracle step failed with code %d" % rc)
        self.t += 1
        return obs, rew

    def reset(self, render=True):
        obs = self._new_obs() if render else None
        rc = lib().orc_reset(self._h, self.B, _p(self.grid), _p(self.pos), _p(self.ori), self.seed,
                             self.env_id_offset, self.t, _p(obs), self.n_threads)
        if rc:
            raise RuntimeError("oracle reset failed with code %d" % rc)
        return obs

    def render(self, rotate=True):
        obs = self._new_obs()
        lib().orc_render(self._h, self.B, _p(self.grid), _p(self.pos), _p(self.ori), int(rotate),
                         _p(obs), self.n_threads)
        return obs
