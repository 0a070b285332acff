#!/usr/bin/env python
"""Sweep the CTA shape of the fused step kernel (SSD_THREADS) on one GPU and
print ms per step for the headline workload.  Usage: python profiles/tune_step.py [game] [B]"""
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequential_social_dilemma_games_b200.batched import BatchedSSDEnv, make_config  # noqa: E402


def time_cfg(game, B, steps=60, warm=10, **env):
    for k, v in env.items():
        os.environ[k] = str(v)
    cfg = make_config(game)
    e = BatchedSSDEnv(cfg, B, seed=0)
    g = torch.Generator(device="cuda").manual_seed(1)
    ring = torch.randint(0, cfg.num_actions, (16, B, cfg.num_agents), generator=g, device="cuda", dtype=torch.int8)
    e.reset()
    for i in range(warm):
        e.step(ring[i % 16])
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(steps):
        e.step(ring[i % 16])
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    e.close()
    return ms


if __name__ == "__main__":
    game = sys.argv[1] if len(sys.argv) > 1 else "harvest"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    for T in (32, 64, 128, 256):
        try:
            ms = time_cfg(game, B, SSD_THREADS=T)
            print("threads=%3d  %.4f ms/step  %.3f G agent-steps/s" % (T, ms, B * 5 / ms / 1e6), flush=True)
        except Exception as ex:
            print("threads=%3d  failed: %s" % (T, ex), flush=True)
