#!/usr/bin/env python
"""Benchmark of the SSD gridworld step path (BASELINE.json: Harvest agent-steps/sec incl. obs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one MapEnv.step over the whole batch of synthetic-action environments resident on a
GPU (one fused kernel launch).  Workload at N=1: BASELINE.json configs[2] -- HarvestEnv, 5 agents,
default HARVEST_MAP, 65 536 batched envs, 15x15x3 egocentric uint8 observations; with N>1 every GPU
owns 65 536 envs of its own (weak scaling, no collective on the step path, Philox streams keyed by
global env id).  Prints ONE JSON line (rank 0).

`value` times K steps of BatchedSSDEnv.step with consecutive steps chained (SSD_OPT_CHAIN_STEPS: programmatic dependent
launch; the actions are pre-generated, which is the option's precondition); `stream_ordered` is the same K steps
without chaining; `e2e` goes through ssd_step_host with pinned host buffers (H2D actions, D2H observations + rewards
inside the timed region); `cpu_baseline` / `--impl reference` time the C port of the reference's step on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "harvest_agent_steps_per_sec_obs_incl"
UNIT = "agent-steps/s"
HORIZON = 1000  # RLlib "horizon" of the reference's training configs (train_baseline.py:131)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=65536)
    ap.add_argument("--game", default="harvest", choices=["harvest", "cleanup"])
    ap.add_argument("--agents", type=int, default=5)
    ap.add_argument("--view", type=int, default=7)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-sample-envs", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-chain", action="store_true", help="stream-ordered steps only (no programmatic dependent launch)")
    return ap.parse_args()


def workload(args):
    return {"workload": "%sEnv %d agents, default map, %d batched envs per GPU, %dx%dx3 uint8 egocentric obs, "
                        "uniform random actions, horizon %d" % (args.game.capitalize(), args.agents, args.envs_per_gpu,
                                                                2 * args.view + 1, 2 * args.view + 1, HORIZON),
            "game": args.game, "num_agents": args.agents, "envs_per_gpu": args.envs_per_gpu,
            "view_radius": args.view, "rng": "philox4x32-10 (production mode)",
            "actions": "pre-generated on device, ring of 16 x [B,N] int8",
            "l2": "per-step working set (state r/w + obs write) is %.0f MB > 126 MB L2; no explicit flush"
                  % ((2 * 608 + 3 * args.agents * (2 * args.view + 1) ** 2) * args.envs_per_gpu / 1e6)}


# ---------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop, self._th = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._th.join(timeout=10)
        return False

    def summary(self):
        sm, reasons, mx = [], set(), None
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_port_rate(args, seconds_target, steps=None, warmup=2):
    """Times oracle/ssd_oracle.c (the CPU port of the reference's step path) on the host cores on a
    bounded sample of the same workload.  Returns (agent-steps/s, cores, sample description, ms/step)."""
    import numpy as np
    from oracle.oracle import OracleEnv
    from sequential_social_dilemma_games_b200.batched import make_config
    cores = min(os.cpu_count() or 1, 256)
    cfg = make_config(args.game, num_agents=args.agents, view_size=args.view)
    B = args.cpu_sample_envs
    env = OracleEnv(cfg, B, seed=0, n_threads=cores)
    env.reset(render=False)
    rng = np.random.RandomState(0)
    ring = rng.randint(cfg.num_actions, size=(16, B, cfg.num_agents)).astype(np.int8)
    buf = env._new_obs()
    for i in range(warmup):
        env.step(ring[i % 16], obs_out=buf)
    n, t0 = 0, time.perf_counter()
    while True:
        env.step(ring[n % 16], obs_out=buf)
        n += 1
        dt = time.perf_counter() - t0
        if (steps is not None and n >= steps) or (steps is None and dt >= seconds_target):
            break
    rate = n * B * cfg.num_agents / dt
    sample = "%d envs x %d steps of the same workload, oracle/ssd_oracle.c, %d pthreads" % (B, n, cores)
    return rate, cores, sample, dt / n * 1e3


def run_reference(args, rank):
    """--impl reference: the reference's step path on the host cores.  The reference is pure Python
    and is not installed on the GPU box; its C port (the oracle, pinned to the reference's golden
    outputs) is what is timed, with all host threads."""
    if rank != 0:
        return
    rate, cores, sample, ms = cpu_port_rate(args, None, steps=args.steps, warmup=max(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": workload(args),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from sequential_social_dilemma_games_b200.batched import BatchedSSDEnv, make_config

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = make_config(args.game, num_agents=args.agents, view_size=args.view)
    B, N = args.envs_per_gpu, cfg.num_agents
    env = BatchedSSDEnv(cfg, B, device=dev, seed=0, env_id_offset=rank * B)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    ring = torch.randint(0, cfg.num_actions, (16, B, N), generator=g, device=dev, dtype=torch.int8)
    obs = torch.empty(env.obs_shape, dtype=torch.uint8, device=dev)
    rew = torch.empty((B, N), dtype=torch.int32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    step_no = [0]

    def one_step():
        if step_no[0] % HORIZON == 0:
            env.reset(out=obs)  # episode boundary, as RLlib's horizon does
        env.step(ring[step_no[0] % 16], out=obs, reward_out=rew)
        step_no[0] += 1

    def timed(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            one_step()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return e0.elapsed_time(e1), float(t.item())

    for _ in range(args.warmup):
        one_step()
    with ClockSampler(local_rank) as clk:  # clocks and throttle reasons over both timed regions
        # stream-ordered steps first (every kernel waits for the previous one to drain) ...
        _, ms_plain = timed(args.steps)
        # ... then the headline: consecutive steps chained with programmatic dependent launch (SSD_OPT_CHAIN_STEPS).
        # The actions are pre-generated, which is the option's precondition; results are identical (tests/test_gpu_parity.py).
        chained = not args.no_chain
        if chained:
            env.chain_steps(True)
            for _ in range(args.warmup):
                one_step()
        launches0 = env.launch_count
        ms, ms_max = timed(args.steps)
    launches = env.launch_count - launches0
    value = world * B * N * args.steps / (ms_max * 1e-3)
    env.chain_steps(False)

    # end to end through the C-ABI with HOST buffers (ssd_step_host): H2D actions, step, D2H obs + rewards
    e2e = None
    if not args.no_e2e:
        a_host = torch.empty((16, B, N), dtype=torch.int8, pin_memory=True)
        a_host.copy_(ring.cpu())
        o_host = torch.empty(env.obs_shape, dtype=torch.uint8, pin_memory=True)
        r_host = torch.empty((B, N), dtype=torch.int32, pin_memory=True)
        a_np, o_np, r_np = a_host.numpy(), o_host.numpy(), r_host.numpy()
        for i in range(3):
            env.step_host(a_np[i % 16], obs_host=o_np, reward_host=r_np)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.e2e_steps):
            env.step_host(a_np[i % 16], obs_host=o_np, reward_host=r_np)
        torch.cuda.synchronize(dev)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * N * args.e2e_steps / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": B * N, "d2h_bytes_per_step": B * (cfg.num_agents * cfg.view_width ** 2 * 3 + 4 * N),
               "api": "ssd_step_host (pinned host buffers: int8 actions in, uint8 obs + int32 rewards out)",
               "steps": args.e2e_steps}

    stats = env.stats()
    tot = torch.tensor([stats["env_steps"], stats["reward_sum"], stats["apples_eaten"], stats["hits"]],
                       dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tot)  # the only collective: end-of-run stats
    clocks = clk.summary()

    # informational: the same step with its consumer on the device (policy forward pass + sampled actions, no PCIe).  Not
    # the contract's e2e; harvest r = 7 only (the trunk kernel is built for 15x15 observations); never fatal.
    on_device = None
    if rank == 0 and args.view == 7 and not args.no_e2e:
        try:
            from sequential_social_dilemma_games_b200 import policy
            net = policy.ConvToFCNet(policy.random_weights(num_outputs=cfg.num_actions, seed=0), device=dev)
            st = {"obs": env.reset(), "hc": net.initial_state(B * N)}

            def loop_step():
                a, _, h, c = net.act(st["obs"].reshape(-1, 15, 15, 3), *st["hc"])
                st["hc"] = (h, c)
                st["obs"], _ = env.step(a.reshape(B, N))
            for _ in range(5):
                loop_step()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                loop_step()
            e1.record()
            torch.cuda.synchronize(dev)
            ms_loop = e0.elapsed_time(e1) / 20
            on_device = {"value": B * N / (ms_loop * 1e-3), "unit": UNIT, "ms_per_step": ms_loop, "n_gpus": 1,
                         "what": "env.step + ConvToFCNet.act (tcgen05 trunk kernel, tcgen05 LSTM + heads + Gumbel-max sampling kernel); "
                                 "observations and actions never leave the device; random-init weights"}
            net.close()
        except Exception as exc:  # noqa: BLE001
            on_device = {"unavailable": repr(exc)[:200]}

    if rank == 0:
        peaks, peak_src = {}, "fallback 6650 GB/s (B200_PROFILING.md)"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
            peak_src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        alg = env.algorithmic_bytes_per_env_step
        kernel_ms = ms / args.steps  # this rank's average step-kernel launch (one launch per step)
        achieved = alg * B / (kernel_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic", "config": dict(workload(args), parallelism="env-sharded x%d" % world,
                                                                    envs_per_cta=env.envs_per_cta,
                                                                    step_launch="chained: programmatic dependent launch, per-warp "
                                                                    "completion words (SSD_OPT_CHAIN_STEPS)" if chained else "stream-ordered"),
                "stream_ordered": {"value": world * B * N * args.steps / (ms_plain * 1e-3), "ms_per_step": ms_plain / args.steps,
                                   "note": "same K steps without chaining: every step kernel drains before the next starts"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "kernel": "ssd_step_fast_kernel<%s,philox,V=%d>" % (args.game.upper(), 2 * args.view + 1), "peak_source": peak_src,
                             "algorithmic_bytes_per_env_step": alg, "units_per_launch": B},
                "clocks": clocks, "gpu_launches": launches, "e2e": e2e,
                "totals": {"env_steps": int(tot[0]), "reward_sum": int(tot[1]), "apples_eaten": int(tot[2]), "hits": int(tot[3])}}
        if on_device is not None:
            line["on_device_loop"] = on_device
        if world == 1 and not args.no_cpu_baseline:
            rate, cores, sample, _ = cpu_port_rate(args, 12.0)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
