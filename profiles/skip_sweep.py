#!/usr/bin/env python
"""Marginal time of the phases of the fused step: SSD_DEBUG_SKIP bits (1 no bulk stores, 2 no render,
4 no spawn, 8 no beams, 16 no literal move emulation, 32 no stats, 64 no grid write-back) at two batch sizes.
Needs a library built with SSD_PROFILING_KNOBS=1 python -m sequential_social_dilemma_games_b200.build --force (a
production build compiles the knobs out).  Results with a skip bit set are WRONG by construction."""
import os
import subprocess
import sys

if len(sys.argv) > 1:
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from tune_step import time_cfg
    for B in [int(x) for x in os.environ.get('BS', '65536,262144').split(',')]:
        ms = time_cfg(os.environ.get("GAME", "harvest"), B, steps=200, warm=100)
        print("skip=%2s B=%6d %.4f ms/step" % (os.environ.get("SSD_DEBUG_SKIP", "0"), B, ms), flush=True)
else:
    for skip in [int(x) for x in os.environ.get('SKIPS', '0,1,2,4,8,6,14').split(',')]:
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, SSD_DEBUG_SKIP=str(skip)))
