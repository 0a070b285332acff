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
