"""Times the fused policy trunk (ssd_policy_features) on the observations of the headline workload:
65 536 Harvest envs x 5 agents = 327 680 agents per call.  Prints ms per call, agents/s, the HBM and the
tensor-core rooflines of the call, and the closed loop env.step -> policy.act -> env.step.
Run on the GPU box: python profiles/policy_bench.py [num_envs]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from sequential_social_dilemma_games_b200 import policy  # noqa: E402
from sequential_social_dilemma_games_b200.batched import BatchedSSDEnv, make_config  # noqa: E402


def timed(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    N = 5
    env = BatchedSSDEnv(make_config("harvest", num_agents=N), B, seed=0)
    net = policy.ConvToFCNet(policy.random_weights(num_outputs=8, seed=0))
    obs = env.reset()
    acts = torch.randint(0, 8, (B, N), dtype=torch.int8, device="cuda")
    for _ in range(20):
        obs, _ = env.step(acts)
    flat = obs.reshape(-1, 15, 15, 3)
    M = flat.shape[0]
    out = torch.empty((M, 32), dtype=torch.float32, device="cuda")
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json"))) \
        if os.path.exists(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")) else {}
    ms = timed(lambda: net.features(flat, out=out), 50)
    bytes_alg = M * (675 + 128)
    flop_useful = M * 2 * (13 * 13 * 6 * 27 + 1014 * 32 + 32 * 32)
    flop_issued = M * 2 * (13 * 80 * 144 + 13 * 32 * 80 + 32 * 32)
    print("trunk: %.4f ms per call, %.3f G agents/s" % (ms, M / ms / 1e6))
    print("  HBM: %.0f GB/s algorithmic (%.1f MB)  peak %s" % (bytes_alg / ms / 1e6, bytes_alg / 1e6, peaks.get("hbm_gbs")))
    print("  tensor: %.1f TFLOP/s useful, %.1f TFLOP/s issued (banded conv)  peak %s" %
          (flop_useful / ms / 1e9, flop_issued / ms / 1e9, peaks.get("bf16_tflops", peaks)))
    h, c = net.initial_state(M)
    ms_fwd = timed(lambda: net.forward(flat, h, c), 20)
    print("forward (trunk + fused LSTM/heads kernel): %.4f ms" % ms_fwd)
    feats = net.features(flat)
    import ctypes as C
    from sequential_social_dilemma_games_b200 import _lib
    hn, cn = torch.empty_like(h), torch.empty_like(c)
    lg = torch.empty((M, 8), device="cuda"); vl = torch.empty((M,), device="cuda"); ac = torch.empty((M,), dtype=torch.int8, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ms_head = timed(lambda: _lib.check(_lib.lib.ssd_policy_lstm_heads(net._h, p(feats), p(h), p(c), p(hn), p(cn), p(lg), p(vl), p(ac), M, 1, 2, st)), 20)
    print("  ssd_policy_lstm_heads alone: %.4f ms (%.0f GB/s of its 2.3 KB per agent)" % (ms_head, M * 2340 / ms_head / 1e6))
    hu, cu = torch.zeros((M, 128), device="cuda"), torch.zeros((M, 128), device="cuda")
    ms_unf = timed(lambda: net.forward_unfused(flat, hu, cu), 20)
    print("forward, unfused route (cuBLAS GEMMs + cell kernel): %.4f ms" % ms_unf)
    state = {"obs": obs, "h": h, "c": c}

    def loop():
        a, _, state["h"], state["c"] = net.act(state["obs"].reshape(-1, 15, 15, 3), state["h"], state["c"])
        state["obs"], _ = env.step(a.reshape(B, N))
    ms_loop = timed(loop, 20)
    print("closed loop env.step + policy.act: %.4f ms per step, %.3f G agent-steps/s" % (ms_loop, M / ms_loop / 1e6))
    ms_env = timed(lambda: env.step(acts), 50)
    print("env.step alone (stream-ordered): %.4f ms" % ms_env)


if __name__ == "__main__":
    main()
