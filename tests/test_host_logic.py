"""Host logic without a GPU: boards, derived point lists, Cleanup probability table, the sharding
plan, and the world_size-2 stats reduction over gloo (the N>1 path's only collective)."""
import hashlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_boards_are_the_reference_boards():
    """SURVEY.md section 0 item 3 (probed on the reference): shapes and point counts; the SHA-256
    values were taken from social_dilemmas/constants.py when the boards were encoded."""
    from sequential_social_dilemma_games_b200.maps import CLEANUP_MAP, HARVEST_MAP, tile_map
    assert (len(HARVEST_MAP), len(HARVEST_MAP[0])) == (16, 38)
    assert (len(CLEANUP_MAP), len(CLEANUP_MAP[0])) == (25, 18)
    hm, cm = "".join(HARVEST_MAP), "".join(CLEANUP_MAP)
    assert (hm.count('A'), hm.count('P'), hm.count('@')) == (155, 20, 104)
    assert (cm.count('B'), cm.count('H'), cm.count('R'), cm.count('S'), cm.count('@'), cm.count('P')) == (103, 56, 63, 12, 82, 10)
    assert hashlib.sha256("\n".join(HARVEST_MAP).encode()).hexdigest() == "e741b7be22f6ec1093f5250b916f59855017dee6102362be91dc3b3655fd0ce4"
    assert hashlib.sha256("\n".join(CLEANUP_MAP).encode()).hexdigest() == "63ae0acace3b26ddd6aa13df7372e1089c5db2fd66db27f8ec157dca380c39e9"
    t = tile_map(CLEANUP_MAP)  # BASELINE.json config 4
    assert (len(t), len(t[0])) == (50, 36) and "".join(t).count('B') == 412 and "".join(t).count('P') == 40


def test_config_tables():
    from sequential_social_dilemma_games_b200.config import (EnvConfig, KIND_CLEANUP, KIND_HARVEST, cleanup_probabilities)
    from sequential_social_dilemma_games_b200.maps import CLEANUP_MAP, HARVEST_MAP
    h = EnvConfig(KIND_HARVEST, HARVEST_MAP, 5)
    assert h.obs_shape == (5, 15, 15, 3) and h.num_actions == 8 and len(h.apple_points) == 155
    assert [tuple(p) for p in h.apple_points[:3]] == [(1, 13), (1, 20), (1, 21)]   # row-major = RNG consumption order
    assert len(h.spawn_points) == 20
    c = EnvConfig(KIND_CLEANUP, CLEANUP_MAP, 5)
    assert c.potential_waste_area == 119 and c.num_actions == 9      # tests/test_envs.py:1009
    assert len(c.spawn_points) == 20                                 # every 'P' twice (cleanup.py:51-52)
    # cleanup.py:156-171 on the default board: no spawning at the initial 56 'H'; h = 47 is the first
    # count below the depletion threshold
    assert c.cleanup_apple_prob[56] == 0 and c.cleanup_waste_prob[56] == 0
    assert c.cleanup_waste_prob[48] == 0 and c.cleanup_waste_prob[47] == 0.5
    assert c.cleanup_apple_prob[0] == 0.05 and 0 < c.cleanup_apple_prob[47] < 0.001
    # the reference's own expectations (tests/test_envs.py:1149-1182): 0 / 0.5 / 0.025
    assert cleanup_probabilities(2, 4) == (0, 0)
    apple, waste = cleanup_probabilities(1, 5)   # density 0.2 -> half of the apple probability
    assert waste == 0.5 and abs(apple - 0.025) < 1e-15
    assert h.colour_lut[ord('A')].tolist() == [0, 255, 0] and c.colour_lut[ord('H')].tolist() == [99, 156, 194]
    g = c.initial_grid()
    assert (g == ord('H')).sum() == 56 and (g == ord('B')).sum() == 0 and (g == ord('P')).sum() == 0
    with pytest.raises(ValueError):
        EnvConfig(KIND_HARVEST, ["@@@", "@ @@", "@@@"], 1)   # ragged board
    with pytest.raises(ValueError):
        EnvConfig(KIND_HARVEST, HARVEST_MAP, 17)


def test_shard_plan():
    from sequential_social_dilemma_games_b200.sharding import shard_range, weak_range
    for total, world in ((65536, 8), (65536, 1), (10, 4), (3, 8), (0, 2), (1000003, 7)):
        spans = [shard_range(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))       # contiguous, disjoint, ordered
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1
    assert weak_range(65536, 3) == (3 * 65536, 4 * 65536)
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from sequential_social_dilemma_games_b200._names import STAT_NAMES
    from sequential_social_dilemma_games_b200.sharding import max_over_ranks, reduce_stats, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    begin, end = shard_range(1001, world, rank)
    local = {k: 0 for k in STAT_NAMES}
    local["env_steps"] = (end - begin) * 10
    local["apples_eaten"] = sum(range(begin, end))   # any per-env quantity: the sum must not depend on the cut
    local["reward_sum"] = -begin
    tot = reduce_stats(local)
    slow = max_over_ranks(1.0 + rank)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, tot, slow, (begin, end)))


def test_two_rank_stats_reduction_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, t0, s0, span0), (r1, t1, s1, span1) = res
    assert t0 == t1 and s0 == s1 == 2.0
    assert span0 == (0, 501) and span1 == (501, 1001)
    assert t0["env_steps"] == 10010 and t0["apples_eaten"] == sum(range(1001)) and t0["reward_sum"] == -501


def test_policy_ref_hand_vector():
    """oracle/policy_ref.py (the float64 checker of the policy kernels, restating models/conv_to_fcnet_v2.py:36-92 and the
    tf.keras 2.0 LSTMCell) on a vector that can be computed by hand: all pixels 255 -> x = 127/255 =: a; every conv tap
    1/27 -> relu(conv) = a; fc1 = 1/1014, fc2 = 1/32 -> features = a; LSTM kernels 0 and bias (i, f, c~, o) = (0, 0, 20, 0)
    -> i = o = 1/2, tanh(20) = 1 to 1e-17, c' = 1/2, h' = tanh(1/2) / 2; logits column j = j / 128 -> j * h', value = -h'."""
    from oracle import policy_ref
    u, A = 128, 4
    w = {"conv_w": np.full((3, 3, 3, 6), 1 / 27.0), "conv_b": np.zeros(6), "fc1_w": np.full((1014, 32), 1 / 1014.0),
         "fc1_b": np.zeros(32), "fc2_w": np.full((32, 32), 1 / 32.0), "fc2_b": np.zeros(32),
         "lstm_w": np.zeros((32, 4 * u)), "lstm_u": np.zeros((u, 4 * u)),
         "lstm_b": np.concatenate([np.zeros(2 * u), np.full(u, 20.0), np.zeros(u)]),
         "logits_w": np.tile(np.arange(A) / 128.0, (u, 1)), "logits_b": np.zeros(A),
         "value_w": np.full((u, 1), -1 / 128.0), "value_b": np.zeros(1)}
    obs = np.full((2, 15, 15, 3), 255, np.uint8)
    a = 127.0 / 255.0
    assert np.allclose(policy_ref.features(w, obs), a, rtol=0, atol=1e-13)
    logits, value, h, c = policy_ref.forward(w, obs, np.zeros((2, u)), np.zeros((2, u)))
    hh = 0.5 * np.tanh(0.5)                       # 0.23105857863000490
    assert abs(hh - 0.2310585786300049) < 1e-15
    assert np.allclose(c, 0.5, rtol=0, atol=1e-12) and np.allclose(h, hh, rtol=0, atol=1e-12)
    assert np.allclose(logits, np.arange(A) * hh, rtol=0, atol=1e-12) and np.allclose(value, -hh, rtol=0, atol=1e-12)
    # second step from that state, gates now see h through a recurrent kernel: u_f = 1/128 on every unit -> z_f = h' * 1 = hh
    w["lstm_u"][:, u:2 * u] = 1 / 128.0
    _, _, h2, c2 = policy_ref.forward(w, obs, h, c)
    f = 1 / (1 + np.exp(-hh))
    assert np.allclose(c2, f * 0.5 + 0.5, rtol=0, atol=1e-12) and np.allclose(h2, 0.5 * np.tanh(f * 0.5 + 0.5), rtol=0, atol=1e-12)
    # an asymmetric conv check: one tap only -> the conv output is that shifted pixel (kernel layout [kh, kw, in, out])
    w2 = dict(w, conv_w=np.zeros((3, 3, 3, 6)))
    w2["conv_w"][2, 0, 1, 4] = 1.0                # filter 4 reads channel 1 at (i + 2, j + 0)
    rng = np.random.RandomState(0)
    ob = rng.randint(0, 256, size=(1, 15, 15, 3)).astype(np.uint8)
    x = (ob.astype(np.float64) - 128.0) / 255.0
    want = np.zeros((13, 13, 6))
    want[:, :, 4] = np.maximum(x[0, 2:15, 0:13, 1], 0)
    fc_in = want.reshape(-1)                      # keras Flatten order (row, col, filter)
    got = policy_ref.features(dict(w2, fc1_w=np.eye(1014)[:, :32] * 1.0, fc2_w=np.eye(32)), ob)
    assert np.allclose(got[0], fc_in[:32], rtol=0, atol=1e-15)


def test_video_helpers(tmp_path):
    """The frame -> file half of the rollout tool (utility_funcs.py:8-56) needs no GPU: PNG frames and an mp4 whose frame
    count and first pixel survive the round trip through OpenCV."""
    cv2 = pytest.importorskip("cv2")
    from sequential_social_dilemma_games_b200 import video
    frames = [np.full((16, 38, 3), 40 * i, np.uint8) for i in range(6)]
    path = video.make_video_from_rgb_imgs(frames, str(tmp_path), video_name="t", fps=8)
    cap = cv2.VideoCapture(path)
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == 6 and int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)) == 640
    img_dir = tmp_path / "frames"
    img_dir.mkdir()
    for i, f in enumerate(frames):
        video.save_img(f, str(img_dir), "frame%06d.png" % i)
    back = cv2.imread(str(img_dir / "frame000003.png"))
    assert back.shape == (16 * 16, 38 * 16, 3) and int(back[0, 0, 0]) == 120
    path2 = video.make_video_from_image_dir(str(tmp_path), str(img_dir), video_name="u", fps=5)
    assert int(cv2.VideoCapture(path2).get(cv2.CAP_PROP_FRAME_COUNT)) == 6


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the arm the driver launches beside ours) on a tiny sample: one JSON line with the contract's
    keys, the same `config` our arm prints, the C port as its value and -- where a reference tree is reachable -- the
    reference's own Python step timed one process per core; and it never maps the product library."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--cpu-sample-envs", "256", "--pyref-steps", "20"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    sys.path.insert(0, root)
    import bench
    ns = type("A", (), dict(game="harvest", agents=5, envs_per_gpu=65536, view=7, gpus=1))()
    assert line["config"] == bench.workload(ns)
    py = line["cpu_baseline"]["python_reference"]
    assert ("aggregate_agent_steps_per_s" in py and py["cores"] >= 1) or "unavailable" in py
    # the reference arm is CPU only: nothing it imports loads libssd_b200.so
    probe = subprocess.run([sys.executable, "-c",
                            "import sys; sys.path.insert(0, %r); import bench, oracle.oracle; "
                            "from sequential_social_dilemma_games_b200.config import make_config; "
                            "print(any('libssd_b200' in l for l in open('/proc/self/maps')))" % root],
                           capture_output=True, text=True, timeout=300, cwd=root)
    assert probe.stdout.strip() == "False", probe.stdout + probe.stderr[-500:]
