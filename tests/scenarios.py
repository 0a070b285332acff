"""Scripted uses of the per-environment API of social_dilemmas/envs (TEST INFRASTRUCTURE).

Every scenario takes an `api` namespace (MapEnv, HarvestEnv, CleanupEnv, Agent, HarvestAgent,
CleanupAgent, BASE_ACTIONS, ...) and a recorder, drives the environments the way the reference's
own tests/test_envs.py does -- subclasses of MapEnv / Agent, num_agents=0 envs with agents added by
hand, set_pos / update_agent_pos / update_agent_rot, world_map edits, update_map, test_map,
step({}) -- and records what comes back.  tests/golden/make_scenarios.py runs them against the
UNMODIFIED reference (build container) and stores the records in tests/golden/scenarios.npz;
tests/test_adapter_gpu.py runs them against sequential_social_dilemma_games_b200.envs on the GPU
and compares record by record, bit for bit.  The scenarios are written for this repository; the
situations they cover are the ones SURVEY.md section 4 lists for the reference's test-suite.
"""
import random

import numpy as np

EMPTY_7 = ['@@@@@@@', '@     @', '@     @', '@     @', '@     @', '@     @', '@@@@@@@']
TWO_SPAWNS = ['@@@@@@', '@ P  @', '@    @', '@    @', '@   P@', '@@@@@@']
ORCHARD = ['@@@@@@@@', '@ P  A @', '@  AAA @', '@ AAAAA@', '@  AAA @', '@ P A P@', '@@@@@@@@']
SMALL_HARVEST = ['@@@@@@', '@ P  @', '@  AA@', '@  AA@', '@  AP@', '@@@@@@']
RIVER = ['@@@@@@@', '@  P  @', '@HHP  @', '@RH   @', '@H P B@', '@SS BB@', '@@@@@@@']
PROB_MAP = ['@@@@@@', '@    @', '@HHPB@', '@RH B@', '@H PB@', '@@@@@@']

ORIENTS = ['UP', 'RIGHT', 'DOWN', 'LEFT']


class Recorder(object):
    def __init__(self):
        self.items = []

    def __call__(self, name, value):
        v = np.array(value)
        if v.dtype.kind == 'U':
            v = np.vectorize(ord)(v).astype(np.int32) if v.size else v.astype(np.int32)
        self.items.append((name, v))


def _make_dummy(api):
    """A MapEnv without custom behaviour and agents that know only the seven base actions."""
    num = {v: k for k, v in api.BASE_ACTIONS.items()}

    class PlainAgent(api.Agent):
        def get_done(self):
            return False

        def action_map(self, action_number):
            return api.BASE_ACTIONS[action_number]

        def consume(self, char):
            return char

        def hit(self, char):
            pass

    class PlainEnv(api.MapEnv):
        def setup_agents(self):
            grid = self.get_map_with_agents()
            for i in range(self.num_agents):
                agent_id = 'agent-' + str(i)
                point = self.spawn_point()
                rot = self.spawn_rotation()
                self.agents[agent_id] = PlainAgent(agent_id, point, rot, grid, 2, 2)

    return PlainEnv, PlainAgent, num


def _refresh(env):
    grid = env.get_map_with_agents()
    for agent in env.agents.values():
        agent.grid = grid


def _place(env, agent_id, pos, rot=None):
    agent = env.agents[agent_id]
    agent.set_pos(pos)
    agent.grid = env.get_map_with_agents()
    agent.update_agent_pos(pos)
    if rot is not None:
        agent.update_agent_rot(rot)


def _add(env, cls, agent_id, pos, rot, *view):
    env.agents[agent_id] = cls(agent_id, pos, rot, env.get_map_with_agents(), *view)
    _refresh(env)


def _state(rec, tag, env):
    rec(tag + '/world', env.world_map)
    rec(tag + '/test_map', env.test_map)
    rec(tag + '/pos', [a.get_pos() for a in env.agents.values()])
    rec(tag + '/ori', [ORIENTS.index(a.get_orientation()) for a in env.agents.values()])


def _step(rec, tag, env, actions, obs_too=True):
    obs, rew, dones, info = env.step(actions)
    _state(rec, tag, env)
    rec(tag + '/rew', [rew[k] for k in sorted(rew)])
    rec(tag + '/dones', [int(dones[k]) for k in sorted(dones)])
    rec(tag + '/info', [len(info)])
    if obs_too:
        for k in sorted(obs):
            rec(tag + '/obs/' + k, obs[k])
    return obs, rew


# ----------------------------------------------------------------------------------------------
def plain_moves(api, rec):
    PlainEnv, PlainAgent, num = _make_dummy(api)
    np.random.seed(3)
    random.seed(3)
    env = PlainEnv(ascii_map=TWO_SPAWNS, num_agents=1)
    obs = env.reset()
    rec('reset/obs', obs['agent-0'])
    _state(rec, 'reset', env)
    rec('base_map', env.base_map)
    for a in range(7):
        _step(rec, 'act%d' % a, env, {'agent-0': a})
    _step(rec, 'noop', env, {})
    # every move action under every orientation, away from the walls and into them
    for rot in ORIENTS:
        for name in ('MOVE_LEFT', 'MOVE_RIGHT', 'MOVE_UP', 'MOVE_DOWN', 'STAY'):
            for start in ([2, 2], [1, 1], [4, 4]):
                _place(env, 'agent-0', start, rot)
                env.step({'agent-0': num[name]})
                rec('table/%s/%s/%s' % (rot, name, start), env.agents['agent-0'].get_pos())
    for rot in ORIENTS:
        for name in ('TURN_CLOCKWISE', 'TURN_COUNTERCLOCKWISE'):
            _place(env, 'agent-0', [2, 2], rot)
            env.step({'agent-0': num[name]})
            rec('turn/%s/%s' % (rot, name), [ORIENTS.index(env.agents['agent-0'].get_orientation())])


def views(api, rec):
    PlainEnv, PlainAgent, num = _make_dummy(api)
    env = PlainEnv(ascii_map=EMPTY_7, num_agents=0)
    env.reset()
    rec('walls/base', env.base_map)
    rec('walls/world', env.world_map)
    _add(env, PlainAgent, 'agent-0', [3, 3], 'UP', 2, 2)
    for pos in ([3, 3], [1, 1], [1, 5], [5, 1], [5, 5], [2, 3], [3, 4], [4, 3], [1, 3]):
        _place(env, 'agent-0', pos)
        _refresh(env)
        rec('view%s' % pos, env.agents['agent-0'].get_state())
    env.update_map([(2, 2, 'A'), (4, 4, 'A')])
    _refresh(env)
    rec('view_apples', env.agents['agent-0'].get_state())
    rec('with_agents', env.get_map_with_agents())
    rec('colors', env.map_to_colors())
    a = np.arange(48).reshape(4, 4, 3)
    for rot in ORIENTS:
        rec('rotate/' + rot, env.rotate_view(rot, a))


def conflicts(api, rec):
    PlainEnv, PlainAgent, num = _make_dummy(api)
    np.random.seed(11)
    random.seed(11)
    env = PlainEnv(ascii_map=TWO_SPAWNS, num_agents=2)
    env.reset()
    _state(rec, 'reset', env)
    # walking into an agent that does not move, then head-on
    _place(env, 'agent-0', [3, 3], 'UP')
    _place(env, 'agent-1', [3, 4], 'UP')
    _step(rec, 'into0', env, {'agent-0': num['MOVE_DOWN']}, False)
    _step(rec, 'into1', env, {'agent-1': num['MOVE_UP']}, False)
    _step(rec, 'swap', env, {'agent-0': num['MOVE_DOWN'], 'agent-1': num['MOVE_UP']}, False)
    # following a leaving agent
    np.random.seed(1)
    for i in range(12):
        _place(env, 'agent-0', [3, 2], 'UP')
        _place(env, 'agent-1', [3, 3], 'UP')
        _step(rec, 'follow%d' % i, env, {'agent-0': num['MOVE_DOWN'], 'agent-1': num['MOVE_LEFT']}, False)
    # two agents want the same cell: the random priority decides
    wins = 0
    for i in range(60):
        _place(env, 'agent-0', [3, 2], 'UP')
        _place(env, 'agent-1', [3, 4], 'UP')
        env.step({'agent-0': num['MOVE_DOWN'], 'agent-1': num['MOVE_UP']})
        wins += env.agents['agent-0'].get_pos().tolist() == [3, 3]
        rec('contest%d' % i, env.test_map)
    rec('contest/wins', [wins])
    # three agents, one cell; then a blocked chain and a rotation cycle of four
    _add(env, PlainAgent, 'agent-2', [2, 3], 'UP', 2, 2)
    for i in range(40):
        _place(env, 'agent-0', [3, 2], 'UP')
        _place(env, 'agent-1', [3, 4], 'UP')
        _place(env, 'agent-2', [2, 3], 'UP')
        env.step({'agent-0': num['MOVE_DOWN'], 'agent-1': num['MOVE_UP'], 'agent-2': num['MOVE_RIGHT']})
        rec('three%d' % i, [a.get_pos() for a in env.agents.values()])
    _add(env, PlainAgent, 'agent-3', [1, 1], 'UP', 2, 2)
    for order in (['agent-0', 'agent-1', 'agent-2', 'agent-3'], ['agent-3', 'agent-1', 'agent-0', 'agent-2']):
        _place(env, 'agent-0', [2, 2], 'UP')
        _place(env, 'agent-1', [2, 3], 'UP')
        _place(env, 'agent-2', [3, 3], 'UP')
        _place(env, 'agent-3', [3, 2], 'UP')
        moves = {'agent-0': 'MOVE_DOWN', 'agent-1': 'MOVE_RIGHT', 'agent-2': 'MOVE_UP', 'agent-3': 'MOVE_LEFT'}
        _step(rec, 'cycle/%s' % order[0], env, {k: num[moves[k]] for k in order}, False)
        _place(env, 'agent-0', [2, 1], 'UP')
        _place(env, 'agent-1', [2, 2], 'UP')
        _place(env, 'agent-2', [2, 3], 'UP')
        _place(env, 'agent-3', [2, 4], 'UP')
        _step(rec, 'chain/%s' % order[0], env, {k: num['MOVE_DOWN'] for k in order}, False)
        _step(rec, 'chain_back/%s' % order[0], env, {k: num['MOVE_UP'] for k in order}, False)


def harvest_small(api, rec):
    num = {v: k for k, v in api.HARVEST_ACTIONS.items()}
    np.random.seed(5)
    random.seed(5)
    env = api.HarvestEnv(ascii_map=SMALL_HARVEST, num_agents=2)
    obs = env.reset()
    for k in sorted(obs):
        rec('reset/obs/' + k, obs[k])
    _state(rec, 'reset', env)
    rec('action_space', [env.action_space.n])
    # eating: +1, the apple disappears
    _place(env, 'agent-0', [2, 2], 'UP')
    _place(env, 'agent-1', [4, 2], 'UP')
    _step(rec, 'eat', env, {'agent-0': num['MOVE_DOWN'], 'agent-1': num['STAY']})
    # beams in all four directions, hitting and missing
    for rot in ORIENTS:
        _place(env, 'agent-0', [2, 2], rot)
        _place(env, 'agent-1', [2, 4], 'UP')
        _step(rec, 'fire/' + rot, env, {'agent-0': num['FIRE']})
        _step(rec, 'after/' + rot, env, {})
    _place(env, 'agent-0', [3, 1], 'RIGHT')
    _place(env, 'agent-1', [3, 2], 'LEFT')
    _step(rec, 'duel', env, {'agent-0': num['FIRE'], 'agent-1': num['FIRE']})
    _step(rec, 'duel_rev', env, {'agent-1': num['FIRE'], 'agent-0': num['FIRE']})
    # regrowth next to the remaining apples
    np.random.seed(2)
    for t in range(60):
        env.step({'agent-0': num['STAY'], 'agent-1': num['STAY']})
        if t % 6 == 5:
            rec('grow%d' % t, env.world_map)
    try:
        env.step({'agent-0': 8})
        rec('bad_action', [0])
    except KeyError:
        rec('bad_action', [1])


def harvest_manual(api, rec):
    num = {v: k for k, v in api.HARVEST_ACTIONS.items()}
    env = api.HarvestEnv(ascii_map=ORCHARD, num_agents=0)
    env.reset()
    _add(env, api.HarvestAgent, 'agent-0', [1, 1], 'UP', 2)
    _add(env, api.HarvestAgent, 'agent-1', [5, 6], 'DOWN', 3)
    rec('start/world', env.world_map)
    rec('view0', env.agents['agent-0'].get_state())
    rng = np.random.RandomState(9)
    np.random.seed(4)
    for t in range(40):
        acts = {'agent-%d' % i: int(rng.randint(8)) for i in rng.permutation(2) if rng.rand() < 0.9}
        _step(rec, 't%d' % t, env, acts, obs_too=(t % 5 == 0))
    env.agents = {}
    env.world_map[2, 4] = ' '
    env.update_map([(1, 1, 'A')])
    _add(env, api.HarvestAgent, 'agent-0', [3, 3], 'LEFT', 2)
    _step(rec, 'after_clear', env, {'agent-0': num['MOVE_UP']})
    rec('count_apples', [env.count_apples(np.array([['A', ' '], ['A', 'A']]))])


def cleanup_small(api, rec):
    num = {v: k for k, v in api.CLEANUP_ACTIONS.items()}
    np.random.seed(7)
    random.seed(7)
    env = api.CleanupEnv(ascii_map=RIVER, num_agents=2)
    obs = env.reset()
    for k in sorted(obs):
        rec('reset/obs/' + k, obs[k])
    _state(rec, 'reset', env)
    rec('area', [env.potential_waste_area])
    rec('action_space', [env.action_space.n])
    # a cleaning beam stops at the first waste cell; a second agent's beam passes where the first one cleaned
    _place(env, 'agent-0', [2, 4], 'UP')
    _place(env, 'agent-1', [3, 4], 'UP')
    _step(rec, 'clean_one', env, {'agent-0': num['CLEAN']})
    _step(rec, 'clean_two', env, {'agent-0': num['CLEAN'], 'agent-1': num['CLEAN']})
    _step(rec, 'clean_rev', env, {'agent-1': num['CLEAN'], 'agent-0': num['CLEAN']})
    _place(env, 'agent-0', [1, 1], 'DOWN')
    _step(rec, 'clean_down', env, {'agent-0': num['CLEAN']})
    # penalty beams ignore waste and stop at agents
    _place(env, 'agent-0', [3, 5], 'UP')
    _place(env, 'agent-1', [3, 3], 'UP')
    _step(rec, 'fire', env, {'agent-0': num['FIRE'], 'agent-1': num['TURN_CLOCKWISE']})
    rec('probs', [env.current_apple_spawn_prob, env.current_waste_spawn_prob])
    np.random.seed(12)
    random.seed(6)
    for t in range(50):
        _step(rec, 'run%d' % t, env, {'agent-0': num['STAY'], 'agent-1': num['CLEAN'] if t % 7 == 0 else num['STAY']},
              obs_too=(t % 10 == 0))
        rec('run%d/probs' % t, [env.current_apple_spawn_prob, env.current_waste_spawn_prob])


def cleanup_probabilities(api, rec):
    env = api.CleanupEnv(ascii_map=PROB_MAP, num_agents=0)
    env.reset()
    rec('area', [env.potential_waste_area])
    rec('p0', [env.current_apple_spawn_prob, env.current_waste_spawn_prob])
    for cells in ([(2, 1)], [(2, 2)], [(3, 2)], [(4, 1)]):
        env.update_map([(r, c, 'R') for r, c in cells])
        env.compute_probabilities()
        rec('p%s' % cells, [env.current_apple_spawn_prob, env.current_waste_spawn_prob])
    np.random.seed(12)
    random.seed(6)
    for t in range(30):
        env.step({})
        rec('spawn%d' % t, env.world_map)
        rec('spawn%d/p' % t, [env.current_apple_spawn_prob, env.current_waste_spawn_prob])
    _add(env, api.CleanupAgent, 'agent-0', [2, 3], 'LEFT', 2)
    random.seed(2)
    np.random.seed(10)
    for t in range(20):
        _step(rec, 'agent%d' % t, env, {'agent-0': 8 if t % 2 == 0 else 4}, obs_too=(t % 5 == 0))


def default_rollouts(api, rec):
    """The stock environments through the module-level generators, construction and reset included."""
    for name, cls, n_act in (('harvest', api.HarvestEnv, 8), ('cleanup', api.CleanupEnv, 9)):
        np.random.seed(21)
        random.seed(21)
        env = cls(num_agents=5)
        obs = env.reset()
        for k in sorted(obs):
            rec('%s/reset/obs/%s' % (name, k), obs[k])
        _state(rec, name + '/reset', env)
        rng = np.random.RandomState(77)
        for t in range(25):
            acts = {'agent-%d' % i: int(rng.randint(n_act)) for i in range(5)}
            if name == 'cleanup' and t < 15:
                for i in range(5):
                    if rng.rand() < 0.5:
                        acts['agent-%d' % i] = 8
            _step(rec, '%s/t%d' % (name, t), env, acts, obs_too=(t % 6 == 0))
        obs = env.reset()
        rec(name + '/reset2/obs0', obs['agent-0'])
        _state(rec, name + '/reset2', env)


def agent_actions_extras(api, rec):
    np.random.seed(8)
    random.seed(8)
    env = api.HarvestEnv(num_agents=3, return_agent_actions=True)
    obs = env.reset()
    for k in sorted(obs):
        for f in ('curr_obs', 'other_agent_actions', 'visible_agents'):
            rec('reset/%s/%s' % (k, f), obs[k][f])
    for t, acts in enumerate(({'agent-0': 1, 'agent-1': 7, 'agent-2': 3}, {'agent-2': 5, 'agent-0': 0, 'agent-1': 2})):
        obs, rew, dones, info = env.step(acts)
        for k in sorted(obs):
            for f in ('curr_obs', 'other_agent_actions', 'visible_agents'):
                rec('t%d/%s/%s' % (t, k, f), obs[k][f])


SCENARIOS = [plain_moves, views, conflicts, harvest_small, harvest_manual, cleanup_small, cleanup_probabilities,
             default_rollouts, agent_actions_extras]


def run_all(api):
    out = {}
    for fn in SCENARIOS:
        rec = Recorder()
        fn(api, rec)
        out[fn.__name__] = rec.items
    return out
