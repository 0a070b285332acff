"""TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Drives the *unmodified* Python reference at /root/reference (vermashresth/
sequential_social_dilemma_games) so that its behaviour can be recorded into golden
fixtures (tests/golden/make_golden.py) and compared with oracle/ssd_oracle.c.

It only works in the build container: /root/reference does not exist on the GPU box, and
nothing under tests/ -m gpu, smoke() or bench.py imports this module.

What it does:
  * registers stand-in modules for `ray.rllib.env`, `gym.spaces`, `matplotlib.pyplot`
    (map_env.py:7,9; harvest.py:1; cleanup.py:1-2) -- none of them is used on the step path;
  * TapeRecorder: wraps np.random.shuffle / np.random.rand / random.shuffle, which the
    reference looks up through the module on every call (map_env.py:422, harvest.py:101,
    cleanup.py:139,145,150), and logs the resulting move order, the uniform doubles and
    the resulting waste-point order per step ("tape-out" replay);
  * PhiloxIn: replaces the same entry points (plus np.random.randint, map_env.py:666) by
    functions driven by the Philox streams of oracle/philox_ref.py ("Philox-in" replay),
    so the reference and the device generate identical draws, reset included.
"""
import os
import random
import sys
import types

import numpy as np

from . import philox_ref as px

REFERENCE_ROOT = os.environ.get("SSD_REFERENCE_ROOT", "/root/reference")

# Device orientation codes (include/ssd_b200.h): clockwise order so a turn is +-1 mod 4.
ORI_CODE = {"UP": 0, "RIGHT": 1, "DOWN": 2, "LEFT": 3}
ORI_NAME = {v: k for k, v in ORI_CODE.items()}


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "social_dilemmas"))


def _install_stubs():
    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    class MultiAgentEnv(object):
        pass

    class _Space(object):
        def __init__(self, *args, **kwargs):
            self.args, self.kwargs = args, kwargs
            if args:
                self.n = args[0]

    if "ray" not in sys.modules:
        env = mod("ray.rllib.env", MultiAgentEnv=MultiAgentEnv)
        rllib = mod("ray.rllib", env=env)
        mod("ray", rllib=rllib)
    if "gym" not in sys.modules:
        spaces = mod("gym.spaces", Box=_Space, Dict=_Space, Discrete=_Space)
        mod("gym", spaces=spaces)
    if "matplotlib" not in sys.modules:
        plt = mod("matplotlib.pyplot")
        mod("matplotlib", pyplot=plt)


_loaded = {}


def load_reference():
    """Import the reference modules; returns a namespace with HarvestEnv, CleanupEnv, ..."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from social_dilemmas.envs import map_env, agent, harvest, cleanup
    from social_dilemmas import constants
    _loaded.update(map_env=map_env, agent=agent, harvest=harvest, cleanup=cleanup,
                   constants=constants, HarvestEnv=harvest.HarvestEnv,
                   CleanupEnv=cleanup.CleanupEnv, MapEnv=map_env.MapEnv)
    return types.SimpleNamespace(**_loaded)


# ----------------------------------------------------------------------------- state I/O
def grid_u8(world_map):
    """<U1 char grid -> uint8 ASCII codes (the C-ABI grid encoding)."""
    return np.vectorize(ord)(world_map).astype(np.uint8)


def extract_state(env):
    """(grid u8[H,W], pos i16[N,2], ori u8[N]) of a reference env."""
    agents = list(env.agents.values())
    pos = np.array([a.pos for a in agents], dtype=np.int16).reshape(len(agents), 2)
    ori = np.array([ORI_CODE[a.orientation] for a in agents], dtype=np.uint8)
    return grid_u8(env.world_map), pos, ori


def obs_u8(obs_f64):
    """Invert map_env.py:199 `(rgb - 128.0) / 255.0` exactly (256 distinct values)."""
    v = np.rint(np.asarray(obs_f64) * 255.0 + 128.0)
    out = v.astype(np.uint8)
    assert np.array_equal((out.astype(np.int64) - 128.0) / 255.0, obs_f64), "obs not on the 256-value lattice"
    return out


# ----------------------------------------------------------------------------- tape-out
class TapeRecorder(object):
    """Context manager recording the reference's RNG results, one record per `begin()`."""

    def __init__(self):
        self._orig = None
        self.begin()

    def begin(self):
        self.move_order = None      # agent indices after np.random.shuffle (map_env.py:422)
        self.uniforms = []          # every np.random.rand(1)[0], in call order
        self.waste_order = None     # env.waste_points after random.shuffle (cleanup.py:145)
        self.spawn_shuffles = 0     # random.shuffle calls on spawn points (map_env.py:656)

    def __enter__(self):
        self._orig = (np.random.shuffle, np.random.rand, random.shuffle)
        o_shuffle, o_rand, o_pyshuffle = self._orig
        rec = self

        def shuffle(x):
            o_shuffle(x)
            # map_env.py:421: list of (agent_id, [row, col]) tuples
            rec.move_order = [int(e[0].split("-")[-1]) for e in x]

        def rand(*shape):
            r = o_rand(*shape)
            rec.uniforms.extend(np.asarray(r, dtype=np.float64).ravel().tolist())
            return r

        def pyshuffle(x, *a):
            o_pyshuffle(x, *a)
            if rec.in_step:
                rec.waste_order = [list(map(int, p)) for p in x]
            else:
                rec.spawn_shuffles += 1

        np.random.shuffle, np.random.rand, random.shuffle = shuffle, rand, pyshuffle
        self.in_step = False
        return self

    def __exit__(self, *exc):
        np.random.shuffle, np.random.rand, random.shuffle = self._orig
        return False

    def step(self, env, actions):
        """env.step(actions) with a fresh record; returns the reference's 4-tuple."""
        self.begin()
        self.in_step = True
        try:
            return env.step(actions)
        finally:
            self.in_step = False


# ----------------------------------------------------------------------------- Philox-in
class PhiloxIn(object):
    """Context manager: the reference draws from the production Philox streams.

    Replacement semantics (each is a valid implementation of the call it replaces):
      np.random.shuffle(x)   Fisher-Yates from the end, j = mulhi32(word, i+1)  (STREAM_MOVE)
      np.random.rand(1)      [u53 / 2**53] with sequential rank k             (STREAM_SPAWN / RSPAWN)
      random.shuffle(x)      x <- sorted(x) ordered by (32-bit key, index)     (STREAM_WASTE / RPOINT)
      np.random.randint(4)   word & 3                                         (STREAM_RROT)
    """

    def __init__(self, seed):
        self.seed = seed
        self.env_id = 0
        self.t = 0
        self.phase = "reset"
        self._orig = None
        self._arm()

    def _arm(self):
        self.k_uniform = 0
        self.k_agent_point = 0
        self.k_agent_rot = 0
        self.n_draws = 0

    def set(self, env_id, t, phase):
        self.env_id, self.t, self.phase = env_id, t, phase
        self._arm()

    def _stream(self, s):
        return px.Stream(self.seed, self.env_id, self.t, s)

    def __enter__(self):
        self._orig = (np.random.shuffle, np.random.rand, random.shuffle, np.random.randint)
        me = self

        def shuffle(x):
            st = me._stream(px.STREAM_MOVE)
            d = 0
            for i in range(len(x) - 1, 0, -1):
                j = (st.word(d) * (i + 1)) >> 32
                d += 1
                x[i], x[j] = x[j], x[i]

        def rand(*shape):
            assert shape == (1,), shape
            st = me._stream(px.STREAM_SPAWN if me.phase == "step" else px.STREAM_RSPAWN)
            u = st.uniform(me.k_uniform)
            me.k_uniform += 1
            return np.array([u], dtype=np.float64)

        def pyshuffle(x, *a):
            canon = sorted([list(map(int, p)) for p in x])
            if me.phase == "step":
                st, base = me._stream(px.STREAM_WASTE), 0
            else:
                st, base = me._stream(px.STREAM_RPOINT), me.k_agent_point * len(canon)
                me.k_agent_point += 1
            keyed = sorted((st.word(base + i), i) for i in range(len(canon)))
            x[:] = [canon[i] for _, i in keyed]

        def randint(n, *a, **k):
            assert n == 4 and not a and not k
            w = me._stream(px.STREAM_RROT).word(me.k_agent_rot)
            me.k_agent_rot += 1
            return int(w & 3)

        np.random.shuffle, np.random.rand, random.shuffle, np.random.randint = \
            shuffle, rand, pyshuffle, randint
        return self

    def __exit__(self, *exc):
        np.random.shuffle, np.random.rand, random.shuffle, np.random.randint = self._orig
        return False
