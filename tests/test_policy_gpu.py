"""The fused policy trunk (csrc/ssd_policy.cu) against a plain PyTorch fp32 reference of the same layers
(models/conv_to_fcnet_v2.py:36-66 on (obs - 128) / 255).  The kernel multiplies fp16 operands (weights and activations
rounded to 11 significant bits) and accumulates in fp32, so the comparison is to a stated tolerance, not bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ATOL, RTOL = 2e-2, 2e-2   # fp16 operand rounding through three layers; features are O(1)


def _reference_features(w, obs):
    x = (obs.to(torch.float32) - 128.0) / 255.0                      # map_env.py:199
    x = x.permute(0, 3, 1, 2)                                         # NHWC -> NCHW
    k = torch.from_numpy(w["conv_w"]).to(obs.device).permute(3, 2, 0, 1)   # [kh, kw, in, out] -> [out, in, kh, kw]
    y = torch.relu(torch.nn.functional.conv2d(x, k, torch.from_numpy(w["conv_b"]).to(obs.device)))
    y = y.permute(0, 2, 3, 1).reshape(obs.shape[0], -1)              # keras Flatten of NHWC
    y = torch.relu(y @ torch.from_numpy(w["fc1_w"]).to(obs.device) + torch.from_numpy(w["fc1_b"]).to(obs.device))
    return torch.relu(y @ torch.from_numpy(w["fc2_w"]).to(obs.device) + torch.from_numpy(w["fc2_b"]).to(obs.device))


@pytest.mark.parametrize("m", [1, 127, 128, 129, 5 * 1024 + 3])
def test_trunk_matches_fp32_reference_random_pixels(m):
    from sequential_social_dilemma_games_b200 import policy
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    w = policy.random_weights(num_outputs=8, seed=3)
    net = policy.ConvToFCNet(w)
    g = torch.Generator(device="cuda").manual_seed(m)
    obs = torch.randint(0, 256, (m, 15, 15, 3), dtype=torch.uint8, device="cuda", generator=g)
    got = net.features(obs)
    want = _reference_features(w, obs)
    torch.cuda.synchronize()
    assert got.shape == (m, 32)
    assert torch.isfinite(got).all()
    err = (got - want).abs().max().item()
    assert torch.allclose(got, want, atol=ATOL, rtol=RTOL), "max abs err %g (ref max %g)" % (err, want.abs().max().item())
    assert want.abs().max().item() > 0.1   # the comparison is not vacuous


def _reference_forward(w, obs, h, c):
    """conv_to_fcnet_v2.py:68-92 in fp32: Keras LSTM cell (gates i, f, c~, o; sigmoid recurrent activation), linear heads."""
    t = lambda k: torch.from_numpy(w[k]).to(obs.device)
    x = _reference_features(w, obs)
    gates = x @ t("lstm_w") + h @ t("lstm_u") + t("lstm_b")
    i, f, g, o = gates.chunk(4, dim=1)
    c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h = torch.sigmoid(o) * torch.tanh(c)
    return h @ t("logits_w") + t("logits_b"), (h @ t("value_w") + t("value_b")).squeeze(1), h, c


def test_forward_matches_fp32_reference_over_steps():
    """Three recurrent steps: bf16 GEMM operands and the fused cell update against the fp32 reference carried alongside."""
    from sequential_social_dilemma_games_b200 import policy
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    m = 1000
    w = policy.random_weights(num_outputs=9, cell_size=128, seed=11)
    net = policy.ConvToFCNet(w)
    g = torch.Generator(device="cuda").manual_seed(5)
    h, c = net.initial_state(m)                                 # tiled layout of the fused kernel
    assert h.shape == (8, 8, 128, 16)
    hr, cr = torch.zeros((m, 128), device="cuda"), torch.zeros((m, 128), device="cuda")
    hu, cu = hr.clone(), cr.clone()
    assert torch.equal(net.state_rows(net.state_from_rows(torch.arange(m * 128.0, device="cuda").reshape(m, 128)), m),
                       torch.arange(m * 128.0, device="cuda").reshape(m, 128))
    for _ in range(3):
        obs = torch.randint(0, 256, (m, 15, 15, 3), dtype=torch.uint8, device="cuda", generator=g)
        logits, value, h, c = net.forward(obs, h, c)            # fused tcgen05 LSTM + heads (cell_size 128)
        l2, v2, h2, c2 = net.forward_unfused(obs, hu, cu)       # library GEMMs + one-pass cell update
        hu, cu = h2, c2
        lr, vr, hr, cr = _reference_forward(w, obs, hr, cr)
        assert logits.shape == (m, 9) and value.shape == (m,)
        for got, want in ((logits, lr), (value, vr), (net.state_rows(h, m), hr), (net.state_rows(c, m), cr)):
            assert torch.allclose(got, want, atol=2e-2, rtol=2e-2), (got - want).abs().max().item()   # fp16 operands, fp32 accumulation
        for got, want in ((l2, lr), (v2, vr), (h2, hr), (c2, cr)):
            assert torch.allclose(got, want, atol=3e-2, rtol=3e-2), (got - want).abs().max().item()   # bf16 operands: 8 significant bits
    net.close()


def test_fused_sampling_follows_softmax():
    """`act` samples inside the kernel (Gumbel-max on Philox): frequencies over many agents with IDENTICAL inputs must match
    softmax(logits); a different counter gives different draws, the same (seed, counter) the same ones."""
    from sequential_social_dilemma_games_b200 import policy
    m = 200000
    w = policy.random_weights(num_outputs=8, cell_size=128, seed=2)
    w["logits_w"] = (w["logits_w"] * 4).astype(np.float32)      # spread the probabilities
    net = policy.ConvToFCNet(w)
    one = torch.randint(0, 256, (1, 15, 15, 3), dtype=torch.uint8, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    obs = one.expand(m, 15, 15, 3).contiguous()
    h, c = net.initial_state(m)
    net.seed_sampling(77)
    a1, value, h1, c1 = net.act(obs, h, c)
    a2, _, _, _ = net.act(obs, h, c)
    net.seed_sampling(77)
    a3, _, _, _ = net.act(obs, h, c)
    logits = net.forward(obs[:1], *net.initial_state(1))[0][0]
    p = torch.softmax(logits.double(), 0).cpu().numpy()
    assert a1.dtype == torch.int8 and int(a1.min()) >= 0 and int(a1.max()) < 8
    freq = np.bincount(a1.cpu().numpy().astype(np.int64), minlength=8) / m
    assert np.abs(freq - p).max() < 5 * np.sqrt(p.max() / m) + 1e-3, (freq, p)   # five sigma of a binomial frequency
    assert torch.equal(a1, a3) and not torch.equal(a1, a2)
    net.close()


def test_trunk_on_env_observations_and_rollout_loop():
    """Observations straight from the step kernel; a short closed loop env -> policy -> env with no host round trip."""
    from sequential_social_dilemma_games_b200 import policy
    from sequential_social_dilemma_games_b200.batched import BatchedSSDEnv, make_config
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    env = BatchedSSDEnv(make_config("harvest", num_agents=5), 300, seed=5)
    w = policy.random_weights(num_outputs=8, seed=1)
    net = policy.ConvToFCNet(w)
    obs = env.reset()
    h, c = net.initial_state(300 * 5)
    gen = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(6):
        flat = obs.reshape(-1, 15, 15, 3)
        got = net.features(flat)
        want = _reference_features(w, flat)
        assert torch.allclose(got, want, atol=ATOL, rtol=RTOL), (got - want).abs().max().item()
        a, value, h, c = net.act(flat, h, c, generator=gen)
        assert a.dtype == torch.int8 and int(a.min()) >= 0 and int(a.max()) < 8
        obs, rew = env.step(a.reshape(300, 5))
    torch.cuda.synchronize()
    env.close()
    net.close()
