#!/usr/bin/env python
"""Wall cycles per phase of one warp task (4 envs) of the specialised step kernel, at several batch sizes -- a lightly loaded GPU
shows the length of the dependency chains, a full one what the phases cost when 32 warps share an SM.
Needs a profiling build: SSD_PROFILING_KNOBS=1 python -m sequential_social_dilemma_games_b200.build --force; the library prints
the averages (SSD_PROF lines on stderr) when a handle is destroyed.
    python profiles/phase_clocks.py [harvest|cleanup] [batch sizes...]
t0 wait for previous kernel, t1 issue loads + zero frames, t2 TMA wait, t3 moves, t4 consume, t5 beams, t6 spawn, t7 grid write-back,
t8 overlay + view params, t9 rows, t10 agent words / rewards / publish."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequential_social_dilemma_games_b200.batched import BatchedSSDEnv  # noqa: E402
from sequential_social_dilemma_games_b200.config import make_config  # noqa: E402

os.environ.pop("SSD_PROF", None)  # only the measured handle below is instrumented
game = sys.argv[1] if len(sys.argv) > 1 else "harvest"
sizes = [int(x) for x in sys.argv[2:]] or [4736, 18944, 65536]
for B in sizes:
    cfg = make_config(game)
    dev = torch.device("cuda", 0)
    warm = BatchedSSDEnv(cfg, B, device=dev, seed=0)
    g = torch.Generator(device=dev).manual_seed(1)
    ring = torch.randint(0, cfg.num_actions, (8, B, cfg.num_agents), generator=g, device=dev, dtype=torch.int8)
    obs = torch.empty(warm.obs_shape, dtype=torch.uint8, device=dev)
    warm.reset(out=obs)
    for i in range(300):   # reach the steady state of the orchard before the measured handle takes over the state
        warm.step(ring[i % 8], out=obs)
    st = warm.get_state()
    os.environ.pop("SSD_PROF", None)
    torch.cuda.synchronize()
    del warm
    os.environ["SSD_PROF"] = "1"
    env = BatchedSSDEnv(cfg, B, device=dev, seed=0)
    env.set_state(*st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200):
        env.step(ring[i % 8], out=obs)
    e1.record()
    torch.cuda.synchronize()
    sys.stderr.write("B=%d %.4f ms/step (instrumented)\n" % (B, e0.elapsed_time(e1) / 200))
    env.close()
