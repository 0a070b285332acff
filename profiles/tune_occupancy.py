#!/usr/bin/env python
"""How sensitive is the fused step kernel to resident warps per SM?  Pads every warp's shared-memory
region by SSD_EXTRA_SMEM bytes and times the headline workload."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from tune_step import time_cfg  # noqa: E402

if __name__ == "__main__":
    B = 65536
    for T in (64, 128):
        for extra in (0, 1024, 2048, 3072, 4352, 6144, 8192):
            ms = time_cfg("harvest", B, SSD_THREADS=T, SSD_EXTRA_SMEM=extra)
            print("threads=%3d extra_smem/warp=%5d  %.4f ms/step  %.3f G agent-steps/s" % (T, extra, ms, B * 5 / ms / 1e6), flush=True)
