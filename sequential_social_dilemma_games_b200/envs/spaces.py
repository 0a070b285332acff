"""gym.spaces when gym(nasium) is installed, otherwise minimal stand-ins with the same attributes
(the reference only builds Box / Dict / Discrete objects: harvest.py:30-44, cleanup.py:68-82)."""
try:  # pragma: no cover - depends on the host environment
    from gym.spaces import Box, Dict, Discrete  # noqa: F401
except Exception:  # pragma: no cover
    try:
        from gymnasium.spaces import Box, Dict, Discrete  # noqa: F401
    except Exception:
        import numpy as np

        class Discrete(object):
            def __init__(self, n):
                self.n = int(n)
                self.shape = ()
                self.dtype = np.int64

            def contains(self, x):
                return 0 <= int(x) < self.n

            def sample(self):
                return int(np.random.randint(self.n))

            def __repr__(self):
                return "Discrete(%d)" % self.n

        class Box(object):
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.low, self.high, self.shape, self.dtype = low, high, tuple(shape) if shape is not None else None, dtype

            def __repr__(self):
                return "Box(%r, %r, %r, %r)" % (self.low, self.high, self.shape, self.dtype)

        class Dict(object):
            def __init__(self, spaces):
                self.spaces = dict(spaces)

            def __getitem__(self, k):
                return self.spaces[k]

            def __repr__(self):
                return "Dict(%r)" % (self.spaces,)
