"""Reference-shaped per-environment API (drop-in for social_dilemmas.envs), CUDA-backed."""
from .agent import (Agent, BASE_ACTIONS, CLEANUP_ACTIONS, CleanupAgent, HARVEST_ACTIONS,  # noqa: F401
                    HarvestAgent, return_view)
from .cleanup import CleanupEnv  # noqa: F401
from .harvest import HarvestEnv  # noqa: F401
from .map_env import ACTIONS, DEFAULT_COLOURS, MapEnv, ORIENTATIONS  # noqa: F401
from .vector_env import SSDVectorEnv  # noqa: F401
