"""B200-native batched Sequential-Social-Dilemma gridworlds (Harvest, Cleanup).

The step path (MapEnv.step / reset of vermashresth/sequential_social_dilemma_games) runs as
hand-written sm_100a CUDA kernels behind the C-ABI of include/ssd_b200.h.  Importing the
compute layer (`batched`, `envs`) requires libssd_b200.so; there is no CPU fallback.
"""
from .config import EnvConfig, KIND_CLEANUP, KIND_HARVEST, KIND_PLAIN, make_config  # noqa: F401
from .maps import CLEANUP_MAP, HARVEST_MAP, tile_map  # noqa: F401

__all__ = ["EnvConfig", "KIND_HARVEST", "KIND_CLEANUP", "KIND_PLAIN", "HARVEST_MAP", "CLEANUP_MAP", "tile_map",
           "BatchedSSDEnv", "make_config"]


def __getattr__(name):  # lazy: `config`/`maps` stay importable where only the CPU tools run
    if name in ("BatchedSSDEnv", "philox_selftest"):
        from . import batched
        return getattr(batched, name)
    raise AttributeError(name)
