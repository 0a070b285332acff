"""GPU parity: the CUDA path (through the C-ABI, libssd_b200.so) against
  (1) the golden outputs of the unmodified Python reference (tests/golden/*.npz), and
  (2) the CPU oracle (oracle/ssd_oracle.c) on larger seeded batches.
Everything is bit-exact: grids, positions, orientations, rewards, uint8 observations."""
import numpy as np
import pytest
import torch

from golden_util import Fixture, PHILOX64_FIXTURES, PHILOX_FIXTURES, TAPE_FIXTURES

pytestmark = pytest.mark.gpu


def _env(cfg, B, **kw):
    from sequential_social_dilemma_games_b200.batched import BatchedSSDEnv
    return BatchedSSDEnv(cfg, B, device="cuda:0", **kw)


def _state(env):
    g, p, o = env.get_state()
    return g.cpu().numpy(), p.cpu().numpy(), o.cpu().numpy()


def _assert_state(env, grid, pos, ori, tag):
    g, p, o = _state(env)
    assert np.array_equal(p, pos), (tag, "pos")
    assert np.array_equal(o, ori), (tag, "ori")
    assert np.array_equal(g, grid), (tag, "grid")


def _assert_env0(env, grid, pos, ori, tag):
    g, p, o = _state(env)
    assert np.array_equal(p[0], pos), (tag, "pos")
    assert np.array_equal(o[0], ori), (tag, "ori")
    assert np.array_equal(g[0], grid), (tag, "grid")


def test_philox_on_device():
    from sequential_social_dilemma_games_b200.batched import philox_selftest
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert tuple(int(x) for x in philox_selftest(ctr, key)) == want


def _order(fx, t, sel=slice(None), explicit=False):
    """The action order of step t, or None when it is the agent order and the caller does not insist."""
    o = fx["order"][t][sel]
    if not explicit and np.array_equal(o, np.broadcast_to(np.arange(o.shape[-1], dtype=o.dtype), o.shape)):
        return None
    return o


@pytest.mark.parametrize("explicit_order", [False, True], ids=["fast", "general"])
@pytest.mark.parametrize("name", TAPE_FIXTURES)
def test_tape_golden(name, explicit_order):
    """Reference trajectories replayed on the device from the recorded RNG tape, through the specialised kernel (whole
    warps; the order array is dropped where it is the agent order) and through the general kernel (explicit order)."""
    fx = Fixture(name)
    rep = 1 if explicit_order else -(-4 // fx.B)  # the specialised kernel steps whole warps of 4 (2) envs
    sel = np.arange(fx.B * rep) % fx.B
    env = _env(fx.cfg, len(sel)).general_kernel_only(explicit_order)
    env.set_state(fx["init_grid"][sel], fx["init_pos"][sel], fx["init_ori"][sel])
    assert np.array_equal(env.render(rotate=False).cpu().numpy(), fx["init_obs"][sel])  # reset() view, map_env.py:239
    for t in range(fx.T):
        obs, rew = env.step(fx["actions"][t][sel], action_order=_order(fx, t, sel, explicit_order), tape=fx.tape(t, sel))
        _assert_state(env, fx["grid"][t][sel], fx["pos"][t][sel], fx["ori"][t][sel], (name, t))
        assert np.array_equal(rew.cpu().numpy(), fx["reward"][t][sel]), (name, t, "reward")
        assert np.array_equal(obs.cpu().numpy(), fx["obs"][t][sel]), (name, t, "obs")
    st = env.stats()
    assert st["env_steps"] == fx.T * len(sel) and st["reward_sum"] == int(fx["reward"][:, sel].sum())


@pytest.mark.parametrize("explicit_order", [False, True], ids=["fast", "general"])
@pytest.mark.parametrize("name", PHILOX_FIXTURES)
def test_philox_golden(name, explicit_order):
    """Reference driven by the production Philox streams: on-device reset + step, no tape.  Every fixture env has its
    own seed, so it runs as env 0 of its own handle -- of 4 envs for the specialised kernel (whole warps; the other
    three are bystanders with neighbouring env ids), of 1 env for the general kernel."""
    fx = Fixture(name)
    reset_at = list(fx["reset_at"])
    nb = 1 if explicit_order else 4
    for b in range(fx.B):
        env = _env(fx.cfg, nb, seed=int(fx["seeds"][b]), env_id_offset=int(fx["env_ids"][b])).general_kernel_only(explicit_order)
        obs = env.reset()
        _assert_env0(env, fx["init_grid"][b], fx["init_pos"][b], fx["init_ori"][b], (name, b, "reset"))
        assert np.array_equal(obs.cpu().numpy()[0], fx["init_obs"][b])
        for t in range(fx.T):
            if t in reset_at:
                ri = reset_at.index(t)
                obs = env.reset()
                _assert_env0(env, fx["reset_grid"][ri, b], fx["reset_pos"][ri, b], fx["reset_ori"][ri, b], (name, b, t, "reset"))
                assert np.array_equal(obs.cpu().numpy()[0], fx["reset_obs"][ri, b])
            assert env.t == t
            order = _order(fx, t, slice(b, b + 1), explicit_order)
            obs, rew = env.step(np.repeat(fx["actions"][t, b:b + 1], nb, axis=0),
                                action_order=None if order is None else np.repeat(order, nb, axis=0))
            _assert_env0(env, fx["grid"][t, b], fx["pos"][t, b], fx["ori"][t, b], (name, b, t))
            assert np.array_equal(rew.cpu().numpy()[0], fx["reward"][t, b]), (name, b, t)
            if fx.has_obs(t):
                assert np.array_equal(obs.cpu().numpy()[0], fx.obs_at(t)[b]), (name, b, t)


def _replay_tiled(name, B, check_every):
    fx = Fixture(name)
    sel = np.arange(B) % fx.B
    env = _env(fx.cfg, B)
    env.set_state(fx["init_grid"][sel], fx["init_pos"][sel], fx["init_ori"][sel])
    assert np.array_equal(env.render(rotate=False).cpu().numpy(), fx["init_obs"][sel])
    for t in range(fx.T):
        obs, rew = env.step(fx["actions"][t][sel], action_order=_order(fx, t, sel), tape=fx.tape(t, sel))  # agent order: specialised kernel
        if t % check_every == 0 or t == fx.T - 1:
            _assert_state(env, fx["grid"][t][sel], fx["pos"][t][sel], fx["ori"][t][sel], (name, t))
        if fx.has_obs(t):
            assert np.array_equal(obs.cpu().numpy(), fx.obs_at(t)[sel]), (name, t)
        assert np.array_equal(rew.cpu().numpy(), fx["reward"][t][sel]), (name, t)
    st = env.stats()
    assert st["env_steps"] == fx.T * B and st["reward_sum"] == int(fx["reward"][:, sel].sum())
    return fx


def test_headline_config_64_reference_envs_production_rng():
    """BASELINE.json configs[2], the benchmark workload itself, against the UNMODIFIED reference under the production random
    streams: 64 reference HarvestEnvs (5 agents, default map) were driven by the Philox streams the kernels draw from
    (tests/golden/make_golden.py, harvest_philox64: construction, reset(), 200 steps, one reset in mid-episode), every env
    with its own seed and global id.  A handle has one seed, so each fixture env runs as env 0 of its own handle of four
    (the other three are bystanders that make the warp whole for the specialised kernel): on-device reset and every step
    bit-exact -- rewards every step, grid / positions / orientations / observations on the steps the fixture keeps."""
    fx = Fixture("harvest_philox64")
    assert fx.B == 64 and fx.T >= 200 and fx.N == 5
    reset_at = list(fx["reset_at"])
    for b in range(fx.B):
        env = _env(fx.cfg, 4, seed=int(fx["seeds"][b]), env_id_offset=int(fx["env_ids"][b]))
        obs = env.reset()
        _assert_env0(env, fx["init_grid"][b], fx["init_pos"][b], fx["init_ori"][b], (b, "reset"))
        assert np.array_equal(obs.cpu().numpy()[0], fx["init_obs"][b])
        rews = []
        for t in range(fx.T):
            if t in reset_at:
                ri = reset_at.index(t)
                obs = env.reset()
                _assert_env0(env, fx["reset_grid"][ri, b], fx["reset_pos"][ri, b], fx["reset_ori"][ri, b], (b, t, "reset"))
                assert np.array_equal(obs.cpu().numpy()[0], fx["reset_obs"][ri, b])
            obs, rew = env.step(np.repeat(fx["actions"][t, b:b + 1], 4, axis=0))
            rews.append(rew[0].clone())
            if fx.has_obs(t):
                _assert_env0(env, fx["grid"][t, b], fx["pos"][t, b], fx["ori"][t, b], (b, t))
                assert np.array_equal(obs.cpu().numpy()[0], fx.obs_at(t)[b]), (b, t)
        assert np.array_equal(torch.stack(rews).cpu().numpy(), fx["reward"][:, b]), b
        env.close()


def test_config2_cleanup_4096_tape():
    """BASELINE.json configs[1] as SURVEY.md 8d specifies it: CleanupEnv, 5 agents, 4096 batched envs on one GPU, bit-exact
    vs the reference under replayed RNG -- 64 DISTINCT reference envs (seeds 100..163) x 200 steps, CLEAN-biased so that the
    waste and fractional-apple paths execute, tiled over the 4096 slots; rewards every step, state every 5th, observations
    on the steps the fixture keeps (every 5th)."""
    fx = _replay_tiled("cleanup_tape64", 4096, 5)
    assert fx.B == 64 and fx.T >= 200


def test_config4_cleanup10_tiled_64_reference_envs():
    """BASELINE.json configs[3] parity spot-check: 64 reference envs on the 2x2-tiled Cleanup map,
    10 agents (agent-9 renders as '1'), P(FIRE) = P(CLEAN) = 0.25, replayed on the device in 1024 slots."""
    fx = _replay_tiled("cleanup10_tiled_tape64", 1024, 4)
    assert fx.B == 64 and fx.N == 10


def _random_actions(rng, cfg, B, p_clean=0.0):
    a = rng.randint(cfg.num_actions, size=(B, cfg.num_agents)).astype(np.int8)
    if p_clean:
        a[rng.rand(B, cfg.num_agents) < p_clean] = 8
    return a


@pytest.mark.parametrize("game,B,steps,N,amap", [("harvest", 2048, 60, 5, None), ("cleanup", 2048, 90, 5, None),
                                                 ("cleanup", 512, 40, 10, "tiled"), ("harvest", 333, 30, 5, None),
                                                 ("harvest", 1024, 800, 5, None), ("harvest", 258, 400, 8, "dense")])
def test_philox_vs_oracle(game, B, steps, N, amap):
    """Production mode (Philox, device reset) against the CPU oracle on the same seeds.  The long Harvest
    cases exist for the neighbour counts the device caches in the grid bytes: a stale count only shows
    up many steps later, when the cell is empty again and draws with the wrong probability."""
    from oracle.oracle import OracleEnv
    from sequential_social_dilemma_games_b200.batched import make_config
    from sequential_social_dilemma_games_b200.maps import CLEANUP_MAP, tile_map
    dense = ['@@@@@@@@@', '@PPPPAAA@', '@PPAAAAA@', '@AAAPPPA@', '@PPPPPAA@', '@AAAAAPP@', '@@@@@@@@@']
    cfg = make_config(game, num_agents=N, ascii_map=tile_map(CLEANUP_MAP) if amap == "tiled" else (dense if amap == "dense" else None))
    seed, off = 0xC0FFEE1234567, 77777
    env = _env(cfg, B, seed=seed, env_id_offset=off)
    orc = OracleEnv(cfg, B, seed=seed, env_id_offset=off, n_threads=8)
    obs = env.reset()
    oobs = orc.reset()
    _assert_state(env, orc.grid, orc.pos, orc.ori, (game, "reset"))
    assert np.array_equal(obs.cpu().numpy(), oobs)
    rng = np.random.RandomState(5)
    for t in range(steps):
        a = _random_actions(rng, cfg, B, p_clean=0.4 if game == "cleanup" and t < steps * 2 // 3 else 0.0)
        obs, rew = env.step(a)
        oobs, orew = orc.step(a)
        assert np.array_equal(rew.cpu().numpy(), orew), (game, t, "reward")
        _assert_state(env, orc.grid, orc.pos, orc.ori, (game, t))
        assert np.array_equal(obs.cpu().numpy(), oobs), (game, t, "obs")
    st = env.stats()
    for i, k in enumerate(("env_steps", "reward_sum", "apples_eaten", "fires", "hits", "cleaned", "apples_spawned", "waste_spawned")):
        assert st[k] == int(orc.stats[i]), (k, st, orc.stats)
    if game == "cleanup":
        assert st["waste_spawned"] > 0 and st["cleaned"] > 0


@pytest.mark.parametrize("game", ["harvest", "cleanup"])
def test_chained_steps(game):
    """SSD_OPT_CHAIN_STEPS: kernels of consecutive steps overlap (programmatic dependent launch, per-warp
    completion words).  Same trajectories, observations, rewards and statistics as stream-ordered steps."""
    from sequential_social_dilemma_games_b200.batched import make_config
    cfg = make_config(game)
    B, T = 32768, 120
    envs = [_env(cfg, B, seed=99, env_id_offset=5), _env(cfg, B, seed=99, env_id_offset=5).chain_steps(True)]
    g = torch.Generator(device="cuda").manual_seed(3)
    ring = torch.randint(0, cfg.num_actions, (8, B, cfg.num_agents), generator=g, device="cuda", dtype=torch.int8)
    outs = []
    for env in envs:
        obs = env.reset().clone()
        rew_sum = torch.zeros((B, cfg.num_agents), dtype=torch.int64, device="cuda")
        snaps = []
        for t in range(T):
            o, r = env.step(ring[t % 8])
            if t % 40 == 39:  # reading results is ordinary stream-ordered work: it must see the finished step
                rew_sum += r
                snaps.append((o.clone(), r.clone()))
        torch.cuda.synchronize()
        outs.append((snaps, [x.cpu().numpy() for x in env.get_state()], env.stats()))
    (s0, st0, k0), (s1, st1, k1) = outs
    for (o0, r0), (o1, r1) in zip(s0, s1):
        assert torch.equal(o0, o1) and torch.equal(r0, r1)
    for x, y in zip(st0, st1):
        assert np.array_equal(x, y)
    assert k0 == k1 and k0["env_steps"] == B * T


@pytest.mark.parametrize("game", ["harvest", "cleanup"])
def test_render_map(game):
    """ssd_render_map == map_to_colors(get_map_with_agents()) (map_env.py:280-339) computed on the host from the state."""
    from sequential_social_dilemma_games_b200.batched import make_config
    cfg = make_config(game, num_agents=10)
    env = _env(cfg, 37, seed=4)
    env.reset()
    rng = np.random.RandomState(1)
    for _ in range(5):
        env.step(_random_actions(rng, cfg, 37))
    frames = env.render_map().cpu().numpy()
    g, p, o = _state(env)
    lut = np.asarray(cfg.colour_lut).reshape(128, 3)
    for b in range(37):
        chars = g[b].copy()
        for ag in range(cfg.num_agents):  # str(int(agent_id[-1]) + 1) in a <U1 array: agent-9 shows as '1'
            chars[p[b, ag, 0], p[b, ag, 1]] = ord(str((ag % 10) + 1)[0])
        assert np.array_equal(frames[b], lut[chars]), (game, b)


def _random_map(rng, game, n_agents):
    H, W = int(rng.randint(5, 25)), int(rng.randint(6, 41))
    m = np.full((H, W), ' ', dtype='<U1')
    m[0, :] = m[-1, :] = m[:, 0] = m[:, -1] = '@'
    inner = [(r, c) for r in range(1, H - 1) for c in range(1, W - 1)]
    rng.shuffle(inner)
    n_wall = len(inner) // 10
    for r, c in inner[:n_wall]:
        m[r, c] = '@'
    rest = inner[n_wall:]
    n_spawn = min(len(rest) // 3, n_agents + int(rng.randint(0, 6)))
    for r, c in rest[:n_spawn]:
        m[r, c] = 'P'
    pool = rest[n_spawn:]
    for r, c in pool[:len(pool) * 2 // 3]:
        m[r, c] = rng.choice(['A', ' ']) if game == "harvest" else rng.choice(['B', 'H', 'R', 'S', ' '])
    return [''.join(row) for row in m], n_spawn


@pytest.mark.parametrize("case", range(14))
def test_random_maps_vs_oracle(case):
    """Random wall-enclosed maps (interior walls, scattered spawn / apple / waste / river cells), 1..16 agents, view radius
    2..10 (5, 7 and 10 take the specialised kernel when the batch fills whole warps), both games, against the oracle."""
    from oracle.oracle import OracleEnv
    from sequential_social_dilemma_games_b200.batched import make_config
    rng = np.random.RandomState(1000 + case)
    game = "harvest" if case % 2 == 0 else "cleanup"
    N = int(rng.randint(1, 17))
    while True:
        amap, n_spawn = _random_map(rng, game, N)
        if n_spawn >= N:
            break
    view = int(rng.choice([2, 3, 5, 5, 7, 7, 10, 10]))
    B = int(rng.choice([8, 24, 37, 64]))
    cfg = make_config(game, num_agents=N, view_size=view, ascii_map=amap)
    env = _env(cfg, B, seed=case, env_id_offset=case * 7)
    orc = OracleEnv(cfg, B, seed=case, env_id_offset=case * 7, n_threads=4)
    assert np.array_equal(env.reset().cpu().numpy(), orc.reset()), (case, "reset")
    _assert_state(env, orc.grid, orc.pos, orc.ori, (case, "reset"))
    for t in range(50):
        a = _random_actions(rng, cfg, B, p_clean=0.3 if game == "cleanup" else 0.0)
        a[rng.rand(B, N) < 0.05] = -1
        obs, rew = env.step(a)
        oobs, orew = orc.step(a)
        assert np.array_equal(rew.cpu().numpy(), orew), (case, t, "reward", amap)
        _assert_state(env, orc.grid, orc.pos, orc.ori, (case, t, amap))
        assert np.array_equal(obs.cpu().numpy(), oobs), (case, t, "obs")


def test_reset_needs_spawn_points():
    """'There are not enough spawn points! Check your map?' (map_env.py:661) is raised by reset, not by construction:
    the adapters place hand-made agents with ssd_set_state on maps with fewer 'P' cells than agents."""
    from sequential_social_dilemma_games_b200 import _lib
    from sequential_social_dilemma_games_b200.config import EnvConfig, KIND_HARVEST
    env = _env(EnvConfig(KIND_HARVEST, ["@@@@@@", "@P AA@", "@    @", "@@@@@@"], 3), 4)
    with pytest.raises(_lib.SsdError, match="not enough spawn points"):
        env.reset()
    grid = np.tile(np.vectorize(ord)(np.array([list(r) for r in ["@@@@@@", "@  AA@", "@    @", "@@@@@@"]])).astype(np.uint8), (4, 1, 1))
    pos = np.tile(np.array([[1, 1], [2, 1], [2, 4]], np.int16), (4, 1, 1))
    env.set_state(grid, pos, np.zeros((4, 3), np.uint8))
    obs, rew = env.step(np.full((4, 3), 3, np.int8))  # MOVE_DOWN facing UP = one column to the right
    assert rew.cpu().numpy().tolist() == [[0, 0, 0]] * 4
    obs, rew = env.step(np.full((4, 3), 3, np.int8))
    assert rew.cpu().numpy().tolist() == [[1, 0, 0]] * 4


@pytest.mark.parametrize("game", ["harvest", "cleanup"])
def test_checkpoint_resume(game):
    """State download + (seed, step counter) is a complete checkpoint: a fresh handle resumed from it continues the
    trajectory bit for bit (Philox streams are pure functions of seed, global env id and step)."""
    from sequential_social_dilemma_games_b200.batched import make_config
    cfg = make_config(game)
    B = 4096
    g = torch.Generator(device="cuda").manual_seed(2)
    acts = torch.randint(0, cfg.num_actions, (60, B, cfg.num_agents), generator=g, device="cuda", dtype=torch.int8)
    a = _env(cfg, B, seed=777, env_id_offset=10)
    a.reset()
    for t in range(30):
        a.step(acts[t])
    ckpt = ([x.cpu().numpy().copy() for x in a.get_state()], a.t)
    b = _env(cfg, B, seed=1, env_id_offset=10)
    b.set_state(*ckpt[0])
    b.seed(777, t=ckpt[1])
    for t in range(30, 60):
        oa, ra = a.step(acts[t])
        ob, rb = b.step(acts[t])
        assert torch.equal(ra, rb) and torch.equal(oa, ob), t
    assert all(torch.equal(x, y) for x, y in zip(a.get_state(), b.get_state()))


@pytest.mark.parametrize("game,B,N", [("harvest", 32768, 5), ("cleanup", 8192, 5), ("harvest", 516, 5), ("harvest", 1030, 5), ("cleanup", 2048, 10)])
def test_rollout_equals_steps(game, B, N):
    """ssd_rollout == T calls of ssd_step: rewards of every step, the observation ring, the final state, statistics and step
    counter.  A batch of less than half a wave of CTAs that goes through the specialised kernel entirely is ONE launch in which
    every warp runs all T steps of its envs; 32 768 envs (throughput-bound anyway) and 1030 envs (a tail for the general kernel)
    are one chained launch per step.  Two rollouts in a row continue the same trajectories."""
    from sequential_social_dilemma_games_b200.batched import make_config
    from sequential_social_dilemma_games_b200.maps import CLEANUP_MAP, tile_map
    cfg = make_config(game, num_agents=N, ascii_map=tile_map(CLEANUP_MAP) if N == 10 else None)
    T, R = 24, 3
    g = torch.Generator(device="cuda").manual_seed(8)
    acts = torch.randint(0, cfg.num_actions, (2 * T, B, cfg.num_agents), generator=g, device="cuda", dtype=torch.int8)
    if game == "cleanup":
        acts[torch.rand(acts.shape, generator=g, device="cuda") < 0.3] = 8   # clean below the depletion threshold: spawning runs
    a, b = _env(cfg, B, seed=12), _env(cfg, B, seed=12)
    a.reset(); b.reset()
    for half in range(2):
        n0 = a.launch_count
        ring, rews = a.rollout(acts[half * T:(half + 1) * T], obs_ring=torch.empty((R,) + tuple(a.obs_shape), dtype=torch.uint8, device="cuda"))
        assert a.launch_count - n0 == (1 if B in (8192, 516, 2048) else T)
        for t in range(T):
            obs, rew = b.step(acts[half * T + t])
            assert torch.equal(rew, rews[t]), (half, t)
            if t >= T - R:
                assert torch.equal(obs, ring[t % R]), (half, t)
        assert all(torch.equal(x, y) for x, y in zip(a.get_state(), b.get_state())) and a.stats() == b.stats() and a.t == b.t
    obs_a, rew_a = a.step(acts[0])      # an ordinary step after a rollout
    obs_b, rew_b = b.step(acts[0])
    assert torch.equal(obs_a, obs_b) and torch.equal(rew_a, rew_b)


@pytest.mark.parametrize("B", [32768, 32770])
def test_chain_is_broken_safely(B):
    """Chained stepping interleaved with everything else a caller may do (masked reset, state download / upload,
    render, a step with an explicit action order, a batch that leaves a tail for the general kernel): same results
    as an unchained handle."""
    from sequential_social_dilemma_games_b200.batched import make_config
    cfg = make_config("harvest")
    g = torch.Generator(device="cuda").manual_seed(5)
    ring = torch.randint(0, cfg.num_actions, (8, B, cfg.num_agents), generator=g, device="cuda", dtype=torch.int8)
    mask = (torch.arange(B, device="cuda") % 3 == 0).to(torch.uint8)
    order = torch.argsort(torch.rand((B, cfg.num_agents), generator=g, device="cuda"), dim=1).to(torch.uint8)
    finals = []
    for chained in (False, True):
        env = _env(cfg, B, seed=31).chain_steps(chained)
        env.reset()
        log = []
        for t in range(30):
            env.step(ring[t % 8])
        env.reset(mask=mask)
        for t in range(30, 45):
            env.step(ring[t % 8])
        st = [x.clone() for x in env.get_state()]
        env.set_state(*st)
        for t in range(45, 60):
            env.step(ring[t % 8])
        log.append(env.render(rotate=True).clone())
        env.step(ring[3], action_order=order)
        for t in range(61, 75):
            obs, rew = env.step(ring[t % 8])
        torch.cuda.synchronize()
        finals.append((log[0], obs.clone(), rew.clone(), [x.cpu().numpy() for x in env.get_state()], env.stats()))
    a, b = finals
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    assert all(np.array_equal(x, y) for x, y in zip(a[3], b[3])) and a[4] == b[4]


def test_full_size_episode_vs_oracle():
    """BASELINE.json configs[2] at full size against the CPU oracle: 65536 Harvest envs, chained steps, 300 steps of one
    episode.  Rewards are compared every step, the full state and the observations every 50 steps, the counters at
    the end -- bit for bit."""
    import os
    from oracle.oracle import OracleEnv
    from sequential_social_dilemma_games_b200.batched import make_config
    cfg = make_config("harvest")
    B, T = 65536, 300
    env = _env(cfg, B, seed=2026).chain_steps(True)
    orc = OracleEnv(cfg, B, seed=2026, n_threads=min(os.cpu_count() or 1, 64))
    assert np.array_equal(env.reset().cpu().numpy(), orc.reset())
    g = torch.Generator(device="cuda").manual_seed(11)
    ring = torch.randint(0, cfg.num_actions, (16, B, cfg.num_agents), generator=g, device="cuda", dtype=torch.int8)
    ring_h = ring.cpu().numpy()
    rews = []
    for t0 in range(0, T, 50):
        for t in range(t0, t0 + 50):  # 50 chained launches back to back, then ordinary stream-ordered reads
            obs, rew = env.step(ring[t % 16])
            rews.append(rew.clone())
        for t in range(t0, t0 + 50):
            oobs, orew = orc.step(ring_h[t % 16])
            assert np.array_equal(rews[t].cpu().numpy(), orew), t
        _assert_state(env, orc.grid, orc.pos, orc.ori, ("full", t0))
        assert np.array_equal(obs.cpu().numpy(), oobs), t0
    st = env.stats()
    for i, k in enumerate(("env_steps", "reward_sum", "apples_eaten", "fires", "hits", "cleaned", "apples_spawned", "waste_spawned")):
        assert st[k] == int(orc.stats[i]), (k, st, orc.stats)


def test_full_size_properties():
    """BASELINE.json configs[2] size (65536 Harvest envs): properties that need no oracle.
    (a) shard invariance: two handles of 32768 envs with env_id_offset reproduce the single
    65536-env handle; (b) determinism; (c) the stats counters equal the summed rewards;
    (d) re-rendering the final state reproduces the observations wherever no beam was drawn."""
    from sequential_social_dilemma_games_b200.batched import make_config
    cfg = make_config("harvest")
    B, steps = 65536, 12
    full = _env(cfg, B, seed=9)
    lo = _env(cfg, B // 2, seed=9)
    hi = _env(cfg, B // 2, seed=9, env_id_offset=B // 2)
    o_full = full.reset().clone()
    o_lo, o_hi = lo.reset().clone(), hi.reset().clone()
    assert torch.equal(o_full[:B // 2], o_lo) and torch.equal(o_full[B // 2:], o_hi)
    g = torch.Generator(device="cuda").manual_seed(3)
    rsum = 0
    for t in range(steps):
        a = torch.randint(0, 8, (B, cfg.num_agents), generator=g, device="cuda", dtype=torch.int8)
        of, rf = full.step(a)
        ol, rl = lo.step(a[:B // 2])
        oh, rh = hi.step(a[B // 2:])
        assert torch.equal(of[:B // 2], ol) and torch.equal(of[B // 2:], oh)
        assert torch.equal(rf[:B // 2], rl) and torch.equal(rf[B // 2:], rh)
        rsum += int(rf.sum().item())
    gf, pf, orf = full.get_state()
    gl, pl, orl = lo.get_state()
    assert torch.equal(gf[:B // 2], gl) and torch.equal(pf[:B // 2], pl) and torch.equal(orf[:B // 2], orl)
    st = full.stats()
    assert st["env_steps"] == B * steps and st["reward_sum"] == rsum
    nofire = (a != 7).all(dim=1)
    last = of.clone()
    again = full.render(rotate=True)
    assert torch.equal(last[nofire], again[nofire]) and int(nofire.sum()) > 1000


def test_edge_cases():
    """Tail CTAs (B not a multiple of the CTA tile), B = 1, empty action dicts, masked reset,
    the host-buffer entry point and the phase-split entry point."""
    from oracle.oracle import OracleEnv
    from sequential_social_dilemma_games_b200 import _lib
    from sequential_social_dilemma_games_b200.batched import make_config
    cfg = make_config("cleanup")
    rng = np.random.RandomState(11)
    for B in (1, 17, 50):
        env = _env(cfg, B, seed=4, env_id_offset=B)
        orc = OracleEnv(cfg, B, seed=4, env_id_offset=B)
        assert np.array_equal(env.reset().cpu().numpy(), orc.reset())
        for t in range(25):
            a = _random_actions(rng, cfg, B, p_clean=0.3)
            if t % 5 == 0:
                a[:] = -1  # step({}) : absent agents do nothing (tests/test_envs.py:437-438)
            if t % 2:
                obs, rew = env.step(a)
                obs, rew = obs.cpu().numpy(), rew.cpu().numpy()
            else:  # ssd_step_host: host buffers in, host buffers out
                obs = np.empty(env.obs_shape, np.uint8)
                _, rew = env.step_host(a, obs_host=obs)
            oobs, orew = orc.step(a)
            assert np.array_equal(obs, oobs) and np.array_equal(rew, orew), (B, t)
            _assert_state(env, orc.grid, orc.pos, orc.ori, (B, t))
    # masked reset: only the selected envs change; the others keep their state
    B = 40
    env = _env(cfg, B, seed=6)
    orc = OracleEnv(cfg, B, seed=6)
    env.reset()
    orc.reset()
    for t in range(10):
        a = _random_actions(rng, cfg, B, p_clean=0.3)
        env.step(a)
        orc.step(a)
    mask = (np.arange(B) % 3 == 0).astype(np.uint8)
    before = _state(env)
    sentinel = torch.full(env.obs_shape, 7, dtype=torch.uint8, device="cuda")
    obs = env.reset(mask=mask, out=sentinel).cpu().numpy()
    fresh = OracleEnv(cfg, B, seed=6)
    fresh.t = orc.t
    fobs = fresh.reset()
    g, p, o = _state(env)
    m = mask.astype(bool)
    assert np.array_equal(g[m], fresh.grid[m]) and np.array_equal(p[m], fresh.pos[m]) and np.array_equal(o[m], fresh.ori[m])
    assert np.array_equal(g[~m], before[0][~m]) and np.array_equal(p[~m], before[1][~m])
    assert np.array_equal(obs[m], fobs[m]) and (obs[~m] == 7).all()
    # phase-split step == fused step
    e1 = _env(cfg, 64, seed=8)
    e2 = _env(cfg, 64, seed=8)
    e1.reset()
    e2.reset()
    for t in range(15):
        a = _random_actions(rng, cfg, 64, p_clean=0.3)
        o1, r1 = e1.step(a)
        r2 = torch.zeros((64, cfg.num_agents), dtype=torch.int32, device="cuda")
        e2.step(a, reward_out=r2, render=False, phases=_lib.PHASE_MOVES | _lib.PHASE_CONSUME)
        e2.step(a, reward_out=r2, render=False, phases=_lib.PHASE_BEAMS)
        o2, _ = e2.step(a, reward_out=r2, phases=_lib.PHASE_SPAWN | _lib.PHASE_RENDER)
        assert torch.equal(o1, o2) and torch.equal(r1, r2), t


@pytest.mark.parametrize("game,N,view,B", [("harvest", 1, 7, 64), ("harvest", 3, 5, 256), ("harvest", 8, 10, 128), ("harvest", 9, 7, 130),
                                           ("harvest", 16, 7, 64), ("cleanup", 1, 5, 64), ("cleanup", 8, 7, 256), ("cleanup", 16, 10, 62),
                                           ("plain", 4, 7, 64)])
def test_specialised_equals_general_kernel(game, N, view, B):
    """The two step kernels share their device functions but not their data structures: the specialised one scans orchard
    bitmaps, keeps a running 'H' count, probes all beams at once and (for these batch sizes) runs in its wide-register
    shape; the general one scans apple points byte by byte, recounts, and fires agent by agent.  Same seeds, same actions:
    identical states, rewards, observations and statistics, for agent counts on both sides of the 8 / 16 lanes-per-env
    split, all three packed view sizes and a map without a game (MapEnv alone).  A scripted rollout in the middle runs as
    one launch where the batch is whole warps."""
    from sequential_social_dilemma_games_b200.config import EnvConfig, KIND_PLAIN, make_config
    from sequential_social_dilemma_games_b200.maps import CLEANUP_MAP, HARVEST_MAP, tile_map
    if game == "plain":
        cfg = EnvConfig(KIND_PLAIN, HARVEST_MAP, N, view_size=view)
    else:
        amap = tile_map(CLEANUP_MAP) if (game == "cleanup" and N > 8) else None   # the default Cleanup map has 10 spawn points
        cfg = make_config(game, num_agents=N, view_size=view, ascii_map=amap)
    a, b = _env(cfg, B, seed=21, env_id_offset=9), _env(cfg, B, seed=21, env_id_offset=9).general_kernel_only(True)
    oa, ob = a.reset(), b.reset()
    assert torch.equal(oa, ob)
    rng = np.random.RandomState(N * 100 + view)
    T = 40
    acts = rng.randint(cfg.num_actions, size=(T, B, N)).astype(np.int8)
    if game == "cleanup":
        acts[rng.rand(T, B, N) < 0.35] = 8
    acts[rng.rand(T, B, N) < 0.05] = -1          # absent agents
    for t in range(T):
        if t == 20:
            dev = torch.from_numpy(acts[20:28]).cuda()
            ring, rews = a.rollout(dev)
            for k in range(8):
                ob, rb = b.step(acts[20 + k])
                assert torch.equal(rews[k], rb), (t, k)
            assert torch.equal(ring[0], ob)
            continue
        if 20 < t < 28:
            continue
        oa, ra = a.step(acts[t])
        ob, rb = b.step(acts[t])
        assert torch.equal(ra, rb), t
        assert torch.equal(oa, ob), t
    assert all(torch.equal(x, y) for x, y in zip(a.get_state(), b.get_state()))
    assert a.stats() == b.stats() and a.t == b.t


@pytest.mark.parametrize("game", ["harvest", "cleanup"])
def test_long_run_with_resets_specialised_equals_general(game):
    """600 steps of 4096 envs with everything that touches the incrementally maintained state in between -- masked resets,
    row-list resets, a state download / upload, a short one-launch rollout -- stepped by the specialised kernel (orchard
    bitmaps, running 'H' count) and by the general kernel (which rebuilds them): a bitmap or count that drifted would change
    which cells draw, and the trajectories would part."""
    from sequential_social_dilemma_games_b200.config import make_config
    cfg = make_config(game)
    B, T = 4096, 600
    a, b = _env(cfg, B, seed=77, env_id_offset=3), _env(cfg, B, seed=77, env_id_offset=3).general_kernel_only(True)
    a.reset(); b.reset()
    g = torch.Generator(device="cuda").manual_seed(31)
    ring = torch.randint(0, cfg.num_actions, (32, B, cfg.num_agents), generator=g, device="cuda", dtype=torch.int8)
    if game == "cleanup":
        ring[torch.rand(ring.shape, generator=g, device="cuda") < 0.3] = 8
    mask = (torch.arange(B, device="cuda") % 7 == 0).to(torch.uint8)
    rows = [5, 6, 7, 8, 1000, 4095]
    for t in range(T):
        if t % 150 == 149:
            a.reset(mask=mask); b.reset(mask=mask)
        if t % 100 == 50:
            a.reset_rows(rows); b.reset_rows(rows)
        if t == 300:
            st = a.get_state()
            a.set_state(*st)
        if t == 400:
            _, rews = a.rollout(ring[:8].contiguous())
            for k in range(8):
                _, rb = b.step(ring[k])
                assert torch.equal(rews[k], rb), (t, k)
            continue
        oa, ra = a.step(ring[t % 32])
        ob, rb = b.step(ring[t % 32])
        if t % 25 == 0 or t == T - 1:
            assert torch.equal(ra, rb) and torch.equal(oa, ob), t
    assert all(torch.equal(x, y) for x, y in zip(a.get_state(), b.get_state()))
    sa, sb = a.stats(), b.stats()
    assert sa == sb and sa["apples_spawned"] > 0 and (game == "harvest" or sa["waste_spawned"] > 0)
