"""Names shared by the binding and the host utilities that must import without the CUDA library."""
STAT_NAMES = ("env_steps", "reward_sum", "apples_eaten", "fires", "hits", "cleaned",
              "apples_spawned", "waste_spawned")   # ssd_stats order, include/ssd_b200.h
