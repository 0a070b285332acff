#!/usr/bin/env python
"""ms/step of the headline workload along one 1000-step episode (random actions deplete the orchard, so
the spawn pass sees more eligible points and fewer apple neighbours late in the episode)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequential_social_dilemma_games_b200.batched import BatchedSSDEnv, make_config  # noqa: E402


def curve(game="harvest", B=65536, window=100, total=1000, view=7, agents=5):
    dev = torch.device("cuda", 0)
    cfg = make_config(game, num_agents=agents, view_size=view)
    env = BatchedSSDEnv(cfg, B, device=dev, seed=0)
    if os.environ.get('SSD_CHAIN', '0') not in ('', '0'):
        env.chain_steps(True)
    g = torch.Generator(device=dev).manual_seed(1234)
    ring = torch.randint(0, cfg.num_actions, (16, B, cfg.num_agents), generator=g, device=dev, dtype=torch.int8)
    obs = torch.empty(env.obs_shape, dtype=torch.uint8, device=dev)
    rew = torch.empty((B, cfg.num_agents), dtype=torch.int32, device=dev)
    env.reset(out=obs)
    for i in range(10):
        env.step(ring[i % 16], out=obs, reward_out=rew)
    env.reset(out=obs)
    out = []
    for w0 in range(0, total, window):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(w0, w0 + window):
            env.step(ring[i % 16], out=obs, reward_out=rew)
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / window)
    return out, env.stats()


if __name__ == "__main__":
    game = sys.argv[1] if len(sys.argv) > 1 else "harvest"
    c, st = curve(game)
    print(game, "ms/step per 100-step window:", " ".join("%.4f" % x for x in c))
    print("episode mean %.4f ms/step = %.3f G agent-steps/s" % (sum(c) / len(c), 65536 * 5 / (sum(c) / len(c)) / 1e6), st)
