"""The CPU oracle (oracle/ssd_oracle.c) against the golden outputs of the unmodified Python
reference (tests/golden/*.npz, written by tests/golden/make_golden.py).  This is the pin that
makes the oracle trustworthy; the GPU parity tests then compare the CUDA path with both."""
import numpy as np
import pytest

from golden_util import Fixture, PHILOX64_FIXTURES, PHILOX_FIXTURES, SEEDED_TAPE_FIXTURES, TAPE_FIXTURES
from oracle.oracle import OracleEnv


@pytest.mark.parametrize("name", TAPE_FIXTURES + SEEDED_TAPE_FIXTURES)
def test_oracle_tape_replay(name):
    fx = Fixture(name)
    env = OracleEnv(fx.cfg, fx.B, n_threads=2)
    env.set_state(fx["init_grid"], fx["init_pos"], fx["init_ori"])
    # reset() renders without rotation (map_env.py:239-240)
    assert np.array_equal(env.render(rotate=False), fx["init_obs"])
    for t in range(fx.T):
        obs, rew = env.step(fx["actions"][t], action_order=fx["order"][t], tape=fx.tape(t))
        assert np.array_equal(env.grid, fx["grid"][t]), (name, t, "grid")
        assert np.array_equal(env.pos, fx["pos"][t]), (name, t, "pos")
        assert np.array_equal(env.ori, fx["ori"][t]), (name, t, "ori")
        assert np.array_equal(rew, fx["reward"][t]), (name, t, "reward")
        assert np.array_equal(env.last_n_draws, fx["n_draws"][t]), (name, t, "n_draws")
        if fx.has_obs(t):
            assert np.array_equal(obs, fx.obs_at(t)), (name, t, "obs")
    assert env.stats[0] == fx.T * fx.B
    assert env.stats[1] == int(fx["reward"].sum())


@pytest.mark.parametrize("name", PHILOX_FIXTURES + PHILOX64_FIXTURES)
def test_oracle_philox_replay(name):
    """Reference driven by the production Philox streams: reset + step parity without a tape."""
    fx = Fixture(name)
    reset_at = list(fx["reset_at"])
    for b in range(fx.B):
        env = OracleEnv(fx.cfg, 1, seed=int(fx["seeds"][b]), env_id_offset=int(fx["env_ids"][b]))
        obs = env.reset()
        assert np.array_equal(env.grid[0], fx["init_grid"][b])
        assert np.array_equal(env.pos[0], fx["init_pos"][b])
        assert np.array_equal(env.ori[0], fx["init_ori"][b])
        assert np.array_equal(obs[0], fx["init_obs"][b])
        for t in range(fx.T):
            if t in reset_at:
                ri = reset_at.index(t)
                obs = env.reset()
                assert np.array_equal(env.grid[0], fx["reset_grid"][ri, b])
                assert np.array_equal(env.pos[0], fx["reset_pos"][ri, b])
                assert np.array_equal(env.ori[0], fx["reset_ori"][ri, b])
                assert np.array_equal(obs[0], fx["reset_obs"][ri, b])
            obs, rew = env.step(fx["actions"][t, b:b + 1], action_order=fx["order"][t, b:b + 1])
            assert np.array_equal(env.grid[0], fx["grid"][t, b]), (name, b, t)
            assert np.array_equal(env.pos[0], fx["pos"][t, b]), (name, b, t)
            assert np.array_equal(env.ori[0], fx["ori"][t, b]), (name, b, t)
            assert np.array_equal(rew[0], fx["reward"][t, b]), (name, b, t)
            assert env.last_n_draws[0] == fx["n_draws"][t, b], (name, b, t)
            if fx.has_obs(t):
                assert np.array_equal(obs[0], fx.obs_at(t)[b]), (name, b, t)


def test_fixture_coverage():
    """The fixtures reach the paths worth pinning."""
    c = Fixture("cleanup_tape")
    assert c["waste_shuffled"].sum() > 100            # waste pass executed
    assert (c["reward"] <= -50).any() and (c["reward"] == 1).any()
    d = Fixture("harvest_dense_tape")
    shared = [len({tuple(p) for p in d["pos"][t, b].tolist()}) < d.N for t in range(d.T) for b in range(d.B)]
    assert any(shared)                                 # two agents in one cell (appendix A.2 item 5)
    assert (d["actions"] < 0).any()                    # partial action dicts
    t10 = Fixture("cleanup10_tiled_tape")
    assert t10.N == 10 and t10.cfg.height == 50 and t10.cfg.width == 36
    # config 2 = 64 distinct reference envs x >= 200 steps with the waste / fractional-apple paths executed,
    # config 4 = 64 reference envs on the tiled map with 10 agents (agent-9 renders as '1')
    c64 = Fixture("cleanup_tape64")
    assert c64.B == 64 and c64.T >= 200 and c64["waste_shuffled"].sum() > 1000 and len(set(c64["seeds"].tolist())) == 64
    t64 = Fixture("cleanup10_tiled_tape64")
    assert t64.B == 64 and t64.N == 10 and t64.cfg.height == 50 and t64["waste_shuffled"].sum() > 100


def test_philox_known_answers():
    """Random123 philox4x32-10 known-answer vectors, C oracle and Python restatement."""
    from oracle import philox_ref
    from oracle.oracle import philox
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert philox_ref.philox4x32_10(ctr, key) == want
        assert tuple(int(x) for x in philox(ctr, key)) == want
