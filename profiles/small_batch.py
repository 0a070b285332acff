#!/usr/bin/env python
"""Step time against batch size for small batches, three ways of driving the same kernel: the Python loop of
BatchedSSDEnv.step (stream-ordered and chained) and ssd_rollout (one C call for T steps, chained).  Separates the host cost of a
step call from what the GPU needs."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequential_social_dilemma_games_b200.batched import BatchedSSDEnv  # noqa: E402
from sequential_social_dilemma_games_b200.config import make_config  # noqa: E402

game = sys.argv[1] if len(sys.argv) > 1 else "harvest"
sizes = [int(x) for x in sys.argv[2:]] or [64, 1024, 4096, 8192, 16384, 32768, 65536]
peak = 6457.4
dev = torch.device("cuda", 0)
print("| envs | python loop stream ms | python loop chained ms | ssd_rollout ms | host s/call us | %% of HBM peak (best) |")
print("|---|---|---|---|---|---|")
for B in sizes:
    cfg = make_config(game)
    env = BatchedSSDEnv(cfg, B, device=dev, seed=0)
    T = 400
    g = torch.Generator(device=dev).manual_seed(1)
    acts = torch.randint(0, cfg.num_actions, (T, B, cfg.num_agents), generator=g, device=dev, dtype=torch.int8)
    obs = torch.empty(env.obs_shape, dtype=torch.uint8, device=dev)
    rew = torch.empty((B, cfg.num_agents), dtype=torch.int32, device=dev)
    env.reset(out=obs)
    res = []
    for chain in (False, True):
        env.chain_steps(chain)
        for i in range(50):
            env.step(acts[i], out=obs, reward_out=rew)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(T):
            env.step(acts[i], out=obs, reward_out=rew)
        e1.record()
        host = (time.perf_counter() - t0) / T * 1e6
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / T)
    env.chain_steps(False)
    ring = torch.empty((1,) + tuple(env.obs_shape), dtype=torch.uint8, device=dev)
    rews = torch.empty((T, B, cfg.num_agents), dtype=torch.int32, device=dev)
    env.rollout(acts[:50], ring, rews[:50])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    env.rollout(acts, ring, rews)
    e1.record()
    torch.cuda.synchronize()
    roll = e0.elapsed_time(e1) / T
    best = min(res + [roll])
    frac = env.algorithmic_bytes_per_env_step * B / (best * 1e-3) / 1e9 / peak
    print("| %d | %.4f | %.4f | %.4f | %.1f | %.1f |" % (B, res[0], res[1], roll, host, 100 * frac), flush=True)
    env.close()
