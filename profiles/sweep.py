#!/usr/bin/env python
"""BASELINE.json configs[4] on one GPU: batch 1 Ki .. 1 Mi envs x view radius 5 / 7 / 10 (Harvest, 5 agents),
stream-ordered and chained steps, plus configs[3] (Cleanup, 10 agents, 2x2-tiled map, 16384 envs).
Prints a markdown table: ms/step, G agent-steps/s, algorithmic GB/s and % of the measured HBM peak."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequential_social_dilemma_games_b200.batched import BatchedSSDEnv, make_config  # noqa: E402
from sequential_social_dilemma_games_b200.maps import CLEANUP_MAP, tile_map  # noqa: E402


def run(cfg, B, chain, steps, warm=30, p_beam=0.0):
    dev = torch.device("cuda", 0)
    env = BatchedSSDEnv(cfg, B, device=dev, seed=0)
    env.chain_steps(chain)
    g = torch.Generator(device=dev).manual_seed(1)
    ring = torch.randint(0, cfg.num_actions if not p_beam else 7, (8, B, cfg.num_agents), generator=g, device=dev, dtype=torch.int8)
    if p_beam:  # configs[3]: P(FIRE) = P(CLEAN) = 0.25, the rest uniform over 0..6
        u = torch.rand((8, B, cfg.num_agents), generator=g, device=dev)
        ring[u < p_beam] = 7
        ring[(u >= p_beam) & (u < 2 * p_beam)] = 8
    obs = torch.empty(env.obs_shape, dtype=torch.uint8, device=dev)
    rew = torch.empty((B, cfg.num_agents), dtype=torch.int32, device=dev)
    env.reset(out=obs)
    for i in range(warm):
        env.step(ring[i % 8], out=obs, reward_out=rew)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        env.step(ring[i % 8], out=obs, reward_out=rew)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    alg = env.algorithmic_bytes_per_env_step
    env.close()
    del obs, env
    torch.cuda.empty_cache()
    return ms, alg


if __name__ == "__main__":
    peak = 6457.4
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    print("| workload | envs | r | steps | ms/step | G agent-steps/s | alg. GB/s | %% of %.0f GB/s |" % peak)
    print("|---|---|---|---|---|---|---|---|")
    for r in (5, 7, 10):
        cfg = make_config("harvest", num_agents=5, view_size=r)
        for B in (1024, 4096, 16384, 65536, 262144, 1048576):
            for chain in (False, True):
                steps = 300 if B <= 65536 else (100 if B <= 262144 else 40)
                ms, alg = run(cfg, B, chain, steps)
                gbs = alg * B / (ms * 1e-3) / 1e9
                print("| harvest N=5 | %d | %d | %s | %.4f | %.3f | %.0f | %.1f |" % (B, r, "chained" if chain else "stream", ms,
                                                                                      B * 5 / ms / 1e6, gbs, 100 * gbs / peak), flush=True)
    cfg = make_config("cleanup", num_agents=10, ascii_map=tile_map(CLEANUP_MAP))
    for chain in (False, True):
        ms, alg = run(cfg, 16384, chain, 200, p_beam=0.25)
        gbs = alg * 16384 / (ms * 1e-3) / 1e9
        print("| cleanup N=10 tiled 2x2 | 16384 | 7 | %s | %.4f | %.3f | %.0f | %.1f |" % ("chained" if chain else "stream", ms,
                                                                                         16384 * 10 / ms / 1e6, gbs, 100 * gbs / peak), flush=True)
    cfg = make_config("cleanup", num_agents=5)
    for B in (4096, 65536):
        for chain in (False, True):
            ms, alg = run(cfg, B, chain, 300)
            gbs = alg * B / (ms * 1e-3) / 1e9
            print("| cleanup N=5 | %d | 7 | %s | %.4f | %.3f | %.0f | %.1f |" % (B, "chained" if chain else "stream", ms, B * 5 / ms / 1e6,
                                                                                gbs, 100 * gbs / peak), flush=True)
