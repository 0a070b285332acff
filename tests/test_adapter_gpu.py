"""the per-environment adapters (sequential_social_dilemma_games_b200.envs) against the
unmodified reference, scenario by scenario, record by record, bit for bit.  The expected records were
produced by tests/golden/make_scenarios.py from /root/reference; here the same scripts (tests/scenarios.py)
drive the adapters, whose every phase runs on the GPU through the C ABI."""
import os
import sys
import types

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import scenarios  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scenarios.npz")


def adapter_api():
    from sequential_social_dilemma_games_b200.envs import agent, cleanup, harvest, map_env
    return types.SimpleNamespace(MapEnv=map_env.MapEnv, HarvestEnv=harvest.HarvestEnv, CleanupEnv=cleanup.CleanupEnv,
                                 Agent=agent.Agent, HarvestAgent=agent.HarvestAgent, CleanupAgent=agent.CleanupAgent,
                                 BASE_ACTIONS=agent.BASE_ACTIONS, HARVEST_ACTIONS=agent.HARVEST_ACTIONS,
                                 CLEANUP_ACTIONS=agent.CLEANUP_ACTIONS)


@pytest.mark.parametrize("fn", scenarios.SCENARIOS, ids=lambda f: f.__name__)
def test_scenario(fn):
    z = np.load(GOLDEN)
    names = [str(x) for x in z[fn.__name__ + "|names"]]
    rec = scenarios.Recorder()
    fn(adapter_api(), rec)
    assert [n for n, _ in rec.items] == names
    for i, (name, v) in enumerate(rec.items):
        want = z["%s|%05d" % (fn.__name__, i)]
        assert v.shape == want.shape, (name, v.shape, want.shape)
        assert np.array_equal(v, want), (fn.__name__, i, name, v, want)


def test_fixture_replay_through_the_dict_api():
    """A tape-out golden fixture (the reference on its own MT19937 streams) replayed through HarvestEnv.step with
    the module-level generators seeded as the reference run was: construction, reset and every step."""
    import random
    from golden_util import Fixture
    from sequential_social_dilemma_games_b200.envs.harvest import HarvestEnv
    fx = Fixture("harvest_tape")
    lut = (np.arange(256) - 128.0) / 255.0
    for b in range(2):
        seed = int(fx["seeds"][b])
        np.random.seed(seed)
        random.seed(seed)
        env = HarvestEnv(ascii_map=fx.ascii_map, num_agents=fx.N)
        obs = env.reset()
        for a in range(fx.N):
            assert np.array_equal(obs['agent-%d' % a], lut[fx["init_obs"][b, a]])
        for t in range(40):
            a8, order = fx["actions"][t, b], fx["order"][t, b]
            obs, rew, dones, info = env.step({'agent-%d' % a: int(a8[a]) for a in order if a8[a] >= 0})
            assert [rew['agent-%d' % a] for a in range(fx.N)] == fx["reward"][t, b].tolist(), (b, t)
            assert np.array_equal(np.vectorize(ord)(env.world_map), fx["grid"][t, b]), (b, t)
            for a in range(fx.N):
                assert np.array_equal(obs['agent-%d' % a], lut[fx["obs"][t, b, a]]), (b, t, a)


@pytest.mark.parametrize("game", ["harvest", "cleanup"])
def test_vector_env_against_oracle(game):
    """SSDVectorEnv (RLlib VectorEnv surface over one batched device state, production Philox streams): list-of-dicts
    in, list-of-dicts out, against the CPU oracle on the same seeds -- partial action dicts, dict iteration order as the
    action order, reset_at of single rows, horizon, return_agent_actions extras."""
    from oracle.oracle import OracleEnv
    from sequential_social_dilemma_games_b200.envs import SSDVectorEnv
    B, N, T = 6, 4, 40
    venv = SSDVectorEnv(game, B, num_agents=N, seed=77, horizon=25, return_agent_actions=True, env_id_offset=3)
    orc = OracleEnv(venv.cfg, B, seed=77, env_id_offset=3, n_threads=2)
    lut = (np.arange(256) - 128.0) / 255.0
    obs = venv.vector_reset()
    oobs = orc.reset()
    for b in range(B):
        for i in range(N):
            o = obs[b]['agent-%d' % i]
            assert np.array_equal(o["curr_obs"], lut[oobs[b, i]])
            assert np.array_equal(o["other_agent_actions"], np.zeros(N - 1, np.int64)) and o["visible_agents"].tolist() == [1] * (N - 1)
    rng = np.random.RandomState(0)
    n_act = venv.action_space.n
    for t in range(T):
        acts, a8, order = [], np.full((B, N), -1, np.int8), np.zeros((B, N), np.uint8)
        for b in range(B):
            perm = rng.permutation(N) if b % 2 else np.arange(N)
            present = [int(i) for i in perm if rng.rand() < 0.85]
            d = {}
            for i in present:
                d['agent-%d' % i] = int(rng.randint(n_act))
                a8[b, i] = d['agent-%d' % i]
            acts.append(d)
            order[b] = present + [i for i in range(N) if i not in present]
        obs, rew, dones, infos = venv.vector_step(acts)
        oobs, orew = orc.step(a8, action_order=order)
        for b in range(B):
            assert infos[b] == {} and dones[b]["__all__"] == (venv._t[b] >= 25)
            assert [rew[b]['agent-%d' % i] for i in range(N)] == orew[b].tolist(), (t, b)
            for i in range(N):
                o = obs[b]['agent-%d' % i]
                assert np.array_equal(o["curr_obs"], lut[oobs[b, i]]), (t, b, i)
                others = [acts[b][k] for k in sorted(acts[b]) if k != 'agent-%d' % i]
                assert o["other_agent_actions"].tolist() == others
    # reset_at: only that row changes, its counter restarts, and the returned (un-rotated) views match a fresh render
    before = [x.cpu().numpy().copy() for x in venv.engine.get_state()]
    o2 = venv.reset_at(2)
    after = [x.cpu().numpy() for x in venv.engine.get_state()]
    keep = [b for b in range(B) if b != 2]
    assert all(np.array_equal(x[keep], y[keep]) for x, y in zip(before, after)) and venv._t[2] == 0 and venv._t[1] == T
    fresh = venv.engine.render(rotate=False).cpu().numpy()
    assert all(np.array_equal(o2['agent-%d' % i]["curr_obs"], lut[fresh[2, i]]) for i in range(N))
    venv.close()


@pytest.mark.parametrize("env_name,render_type", [("harvest", "fast"), ("cleanup", "pretty")])
def test_rollout_controller_writes_a_video(tmp_path, env_name, render_type):
    """SURVEY 8f-4 tooling: the reference's rollout.Controller (rollout.py:30-110) on the batched engine -- frames from
    ssd_render_map, video through OpenCV.  The frames are the full map with the agents painted: agent-0's colour is in it."""
    cv2 = pytest.importorskip("cv2")
    from sequential_social_dilemma_games_b200.rollout import Controller
    c = Controller(env_name=env_name, num_envs=3, film=1, seed=4)
    rewards, observations, full_obs = c.rollout(horizon=4)
    assert len(rewards) == len(observations) == len(full_obs) == 4
    assert full_obs[0].shape == (c.cfg.height, c.cfg.width, 3) and full_obs[0].dtype == np.uint8
    assert observations[0].shape == (15, 15, 3) and observations[0].dtype == np.float64 and np.abs(observations[0]).max() <= 0.502
    assert (full_obs[-1] == np.array([159, 67, 255], np.uint8)).all(-1).sum() == 1     # agent-0 ('1'), map_env.py:31
    path = c.render_rollout(horizon=5, path=str(tmp_path), render_type=render_type, fps=8)
    assert path.endswith(env_name + "_trajectory.mp4")
    assert int(cv2.VideoCapture(path).get(cv2.CAP_PROP_FRAME_COUNT)) == 5
    c.close()
