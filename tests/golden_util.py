"""Loading and replaying the golden fixtures under tests/golden (test infrastructure)."""
import os

import numpy as np

from sequential_social_dilemma_games_b200.config import EnvConfig

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

def regenerate_tape(d):
    """Uniforms [T, B, K] and waste orders [T, B, nw] of a "seeded tape" fixture (tests/golden/make_golden.py): the
    reference ran on np.random / random seeded per env; their MT19937 states right after reset() are stored, and one
    step consumes, in this order, np.random.shuffle of the m movers (map_env.py:421-423; m from the recorded move
    order), n_draws np.random.rand values (harvest.py:101, cleanup.py:139,150) and -- when the waste pass ran -- one
    random.shuffle of the persistent waste_points list (cleanup.py:145)."""
    import random
    T, B = d["actions"].shape[:2]
    nd, mo = d["n_draws"], d["move_order"]
    K = max(int(nd.max()), 1)
    nw = d["waste_init"].shape[1] if "waste_init" in d else 0
    u = np.zeros((T, B, K), np.float64)
    w = np.zeros((T, B, nw), np.uint16)
    for b in range(B):
        st = d["np_state"][b]
        rs = np.random.RandomState()
        rs.set_state(('MT19937', st[:624].astype(np.uint32), int(st[624]), 0, 0.0))
        pr = random.Random()
        pr.setstate((3, tuple(int(x) for x in d["py_state"][b]), None))
        waste = [int(x) for x in d["waste_init"][b]] if nw else []
        for t in range(T):
            m = int((mo[t, b] != 255).sum())
            if m:
                rs.shuffle(list(range(m)))
            n = int(nd[t, b])
            if n:
                u[t, b, :n] = rs.random_sample(n)
            if nw:
                if d["waste_shuffled"][t, b]:
                    pr.shuffle(waste)
                w[t, b] = waste
    return u, w


TAPE_FIXTURES = ["harvest_tape", "cleanup_tape", "cleanup10_tiled_tape", "harvest_dense_tape",
                 "harvest_r5_tape", "harvest_r10_tape", "cleanup_order_tape"]
SEEDED_TAPE_FIXTURES = ["cleanup_tape64", "cleanup10_tiled_tape64"]  # 64 reference envs each, tape regenerated from MT19937 states
PHILOX64_FIXTURES = ["harvest_philox64"]  # the headline workload: 64 reference envs driven by the production Philox streams
PHILOX_FIXTURES = ["harvest_philox", "cleanup_philox", "cleanup10_tiled_philox"]


class Fixture(object):
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.d = {k: z[k] for k in z.files}
        self.kind = int(self.d["kind"])
        self.ascii_map = [str(r) for r in self.d["ascii_map"]]
        self.N = int(self.d["num_agents"])
        self.view = int(self.d["view_size"])
        self.mode = str(self.d["mode"])
        self.cfg = EnvConfig(self.kind, self.ascii_map, self.N, view_size=self.view)
        self.T, self.B = self.d["actions"].shape[:2]
        self.obs_steps = set(int(t) for t in self.d["obs_steps"]) if "obs_steps" in self.d else None
        if self.mode == "tape" and "np_state" in self.d:  # seeded tape: regenerate uniforms and waste orders
            self.uniforms, self.d["waste_order"] = regenerate_tape(self.d)
        elif self.mode == "tape":
            nd = self.d["n_draws"]
            K = max(int(nd.max()), 1)
            self.uniforms = np.zeros((self.T, self.B, K), np.float64)
            off = 0
            flat = self.d["u_flat"]
            for t in range(self.T):
                for b in range(self.B):
                    n = int(nd[t, b])
                    self.uniforms[t, b, :n] = flat[off:off + n]
                    off += n
            assert off == len(flat)

    def __getitem__(self, k):
        return self.d[k]

    def has_obs(self, t):
        return self.obs_steps is None or t in self.obs_steps

    def obs_at(self, t):
        """Reference observations after step t (fixtures with obs_steps keep every k-th step only)."""
        if self.obs_steps is None:
            return self.d["obs"][t]
        return self.d["obs"][int(np.searchsorted(self.d["obs_steps"], t))]

    def tape(self, t, envs=None):
        """Tape dict of step t for env rows `envs` (index array; default all)."""
        sel = slice(None) if envs is None else envs
        return dict(move_order=np.ascontiguousarray(self.d["move_order"][t][sel]),
                    uniforms=np.ascontiguousarray(self.uniforms[t][sel]),
                    waste_order=np.ascontiguousarray(self.d["waste_order"][t][sel])
                    if self.d["waste_order"].shape[-1] else None)
