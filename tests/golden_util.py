"""Loading and replaying the golden fixtures under tests/golden (test infrastructure)."""
import os

import numpy as np

from sequential_social_dilemma_games_b200.config import EnvConfig

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TAPE_FIXTURES = ["harvest_tape", "cleanup_tape", "cleanup10_tiled_tape", "harvest_dense_tape",
                 "harvest_r5_tape", "harvest_r10_tape", "cleanup_order_tape"]
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
        if self.mode == "tape":
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

    def tape(self, t, envs=None):
        """Tape dict of step t for env rows `envs` (index array; default all)."""
        sel = slice(None) if envs is None else envs
        return dict(move_order=np.ascontiguousarray(self.d["move_order"][t][sel]),
                    uniforms=np.ascontiguousarray(self.uniforms[t][sel]),
                    waste_order=np.ascontiguousarray(self.d["waste_order"][t][sel])
                    if self.d["waste_order"].shape[-1] else None)
