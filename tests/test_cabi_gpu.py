"""C-ABI behaviour that needs a device: row-list reset, state validation, parked agents, ordering of ssd_step_host
behind asynchronous work, beams that must not survive into a later phase-split step."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _env(cfg, B, **kw):
    from sequential_social_dilemma_games_b200.batched import BatchedSSDEnv
    return BatchedSSDEnv(cfg, B, device="cuda:0", **kw)


def _state(env):
    return tuple(x.cpu().numpy() for x in env.get_state())


@pytest.mark.parametrize("game,N", [("harvest", 5), ("cleanup", 5), ("cleanup", 10)])
def test_reset_rows_equals_masked_reset(game, N):
    """ssd_reset_rows(rows) == ssd_reset(mask) on the same rows: same agents, grids and (un-rotated) observations for the
    listed envs, nothing else touched -- including two rows of one warp's group of envs and a duplicated row."""
    from sequential_social_dilemma_games_b200.config import make_config
    cfg = make_config(game, num_agents=N)
    B = 203
    rows = [0, 1, 7, 64, 65, 130, 202, 7]
    a, b = _env(cfg, B, seed=5, env_id_offset=11), _env(cfg, B, seed=5, env_id_offset=11)
    rng = np.random.RandomState(0)
    for env in (a, b):
        env.reset()
    for t in range(6):
        act = rng.randint(cfg.num_actions, size=(B, N)).astype(np.int8)
        a.step(act)
        b.step(act)
    before = _state(a)
    mask = np.zeros(B, np.uint8)
    mask[rows] = 1
    obs_a = a.reset(mask=mask).cpu().numpy()
    launches = b.launch_count
    obs_b = b.reset_rows(rows).cpu().numpy()
    assert b.launch_count - launches == 2          # one reset launch + one spawn/render launch over len(rows) warps
    sa, sb = _state(a), _state(b)
    for x, y in zip(sa, sb):
        assert np.array_equal(x, y)
    assert np.array_equal(obs_a[rows], obs_b[rows])
    untouched = np.setdiff1d(np.arange(B), rows)
    for x, y in zip(before, sb):
        assert np.array_equal(x[untouched], y[untouched])
    # a device tensor of rows works too; out-of-range host rows are refused
    b.reset_rows(torch.tensor([3, 4], dtype=torch.int32, device="cuda"))
    from sequential_social_dilemma_games_b200 import _lib
    with pytest.raises(_lib.SsdError):
        b.reset_rows([B])


def test_vector_env_reset_at_is_one_row():
    from sequential_social_dilemma_games_b200.envs.vector_env import SSDVectorEnv
    v = SSDVectorEnv("harvest", 40, seed=3, uint8_obs=True)
    v.vector_reset()
    acts = [{"agent-%d" % i: 4 for i in range(5)} for _ in range(40)]
    v.vector_step(acts)
    g0 = _state(v.engine)
    n0 = v.engine.launch_count
    o = v.reset_at(17)
    assert v.engine.launch_count - n0 == 2 and set(o.keys()) == {"agent-%d" % i for i in range(5)}
    g1 = _state(v.engine)
    keep = np.arange(40) != 17
    for x, y in zip(g0, g1):
        assert np.array_equal(x[keep], y[keep])
    v.close()


def test_set_state_rejects_positions_outside_the_map():
    from sequential_social_dilemma_games_b200 import _lib
    from sequential_social_dilemma_games_b200.config import make_config
    cfg = make_config("harvest")
    env = _env(cfg, 8, seed=1)
    env.reset()
    g, p, o = _state(env)
    for bad in ((16, 3), (3, 38), (-1, 3), (3, 300)):
        q = p.copy()
        q[5, 2] = bad
        with pytest.raises(_lib.SsdError, match="outside"):
            env.set_state(g, q, o)
        for x, y in zip((g, p, o), _state(env)):   # the state was not touched
            assert np.array_equal(x, y)
    env.set_state(g, p, o)


def test_agent_on_a_wall_cell_is_parked():
    """An agent uploaded onto '@' (the adapters' stand-in for an env without agents) never acts, is never painted and no
    beam starts from it; the other agents step exactly as the oracle steps them without it firing."""
    from oracle.oracle import OracleEnv
    from sequential_social_dilemma_games_b200.config import make_config
    cfg = make_config("harvest", num_agents=3)
    B = 8
    env, orc = _env(cfg, B, seed=2), OracleEnv(cfg, B, seed=2)
    env.reset()
    g, p, o = _state(env)
    p[:, 2] = (0, 0)          # agent-2 on the corner wall, facing UP: a FIRE from there would leave the map
    o[:, 2] = 0
    env.set_state(g, p, o)
    orc.set_state(g, p, o)
    rng = np.random.RandomState(1)
    for t in range(20):
        act = rng.randint(cfg.num_actions, size=(B, 3)).astype(np.int8)
        act[:, 2] = 7         # the parked agent "fires" every step
        oact = act.copy()
        oact[:, 2] = -1       # the oracle sees it as absent
        obs, rew = env.step(act)
        oobs, orew = orc.step(oact)
        g1, p1, o1 = _state(env)
        assert np.array_equal(p1, orc.pos) and np.array_equal(o1, orc.ori) and np.array_equal(g1, orc.grid), t
        assert np.array_equal(rew.cpu().numpy(), orew), t
        # the oracle paints its (absent but present) agent-2 on the wall cell; the device does not: compare the views of
        # the two real agents wherever the corner is out of sight, and make sure the corner itself stays a wall
        ob = obs.cpu().numpy()
        far = (p1[:, :2, 0] > 8) | (p1[:, :2, 1] > 8)
        assert np.array_equal(ob[:, :2][far], oobs[:, :2][far]), t
    assert (env.stats()["fires"] == orc.stats[3])


def test_step_host_waits_for_the_callers_stream():
    """ssd_step_host runs on internal streams; it must see an asynchronous reset / set_state queued just before it."""
    from oracle.oracle import OracleEnv
    from sequential_social_dilemma_games_b200.config import make_config
    cfg = make_config("harvest")
    B = 16384
    env, orc = _env(cfg, B, seed=9), OracleEnv(cfg, B, seed=9, n_threads=8)
    rng = np.random.RandomState(3)
    act = rng.randint(cfg.num_actions, size=(B, 5)).astype(np.int8)
    obs_h = np.zeros(env.obs_shape, np.uint8)
    rew_h = np.zeros((B, 5), np.int32)
    junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        junk.zero_()           # keep the caller's stream busy so that the reset below is still pending
    env.reset(render=False)    # asynchronous
    env.step_host(act, obs_host=obs_h, reward_host=rew_h)
    orc.reset(render=False)
    oobs, orew = orc.step(act)
    assert np.array_equal(rew_h, orew) and np.array_equal(obs_h, oobs)
    g, p, o = _state(env)
    assert np.array_equal(g, orc.grid) and np.array_equal(p, orc.pos)


def test_phase_split_beams_do_not_leak_into_the_next_step():
    """Step 1 runs the device beam phase (FIRE recorded in the beam buffer, no render).  Step 2 moves and renders without
    a beam phase (a caller whose own custom_action hook ran instead): its observations must show no stale 'F' cells."""
    from sequential_social_dilemma_games_b200 import _lib
    from sequential_social_dilemma_games_b200.config import make_config
    cfg = make_config("harvest")
    B = 4
    env = _env(cfg, B, seed=4)
    env.reset()
    fire = np.full((B, 5), 7, np.int8)
    stay = np.full((B, 5), 4, np.int8)
    env.step(fire, phases=_lib.PHASE_MOVES | _lib.PHASE_CONSUME | _lib.PHASE_BEAMS, render=False)
    assert env.get_beams()[:, 48:53].any()
    env.step(stay, phases=_lib.PHASE_SPAWN, render=False)
    env.step(stay, phases=_lib.PHASE_MOVES | _lib.PHASE_CONSUME, render=False)   # new step, no device beam phase
    obs, _ = env.step(stay, phases=_lib.PHASE_SPAWN | _lib.PHASE_RENDER)
    yellow = (obs.cpu().numpy() == np.array([255, 255, 0], np.uint8)).all(-1)
    assert not yellow.any()
    assert not env.get_beams()[:, 48:53].any()
