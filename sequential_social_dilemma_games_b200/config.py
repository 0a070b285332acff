"""Host-side derivation of everything the C-ABI needs from an ASCII map.

Mirrors what the reference computes in its constructors (map_env.py:62-102,
harvest.py:20-28, cleanup.py:32-66) and the module constants it reads at call time
(harvest.py:8-15, cleanup.py:11-27, map_env.py:24-41).  The Cleanup probability table is
evaluated here with the reference's own Python expression (cleanup.py:156-171) for every
possible waste count, so the device only ever compares doubles -- it never re-derives them.
"""
import numpy as np

from .maps import CLEANUP_MAP, HARVEST_MAP, validate_map

KIND_HARVEST, KIND_CLEANUP, KIND_PLAIN = 0, 1, 2
MAX_AGENTS = 16

# map_env.py:24-41 (DEFAULT_COLOURS) and cleanup.py:15-18 (CLEANUP_COLORS).  The reference
# mutates its module-global dict when the first CleanupEnv is built (cleanup.py:64); both
# games therefore share one table here.
DEFAULT_COLOURS = {
    ' ': (0, 0, 0), '0': (0, 0, 0), '@': (180, 180, 180), 'A': (0, 255, 0), 'F': (255, 255, 0),
    'P': (159, 67, 255),
    '1': (159, 67, 255), '2': (2, 81, 154), '3': (204, 0, 204), '4': (216, 30, 54),
    '5': (254, 151, 0), '6': (100, 255, 255), '7': (99, 99, 255), '8': (250, 204, 255),
    '9': (238, 223, 16),
}
CLEANUP_COLOURS = {'C': (100, 255, 255), 'S': (113, 75, 24), 'H': (99, 156, 194), 'R': (113, 75, 24)}

HARVEST_SPAWN_PROB = (0, 0.005, 0.02, 0.05)   # harvest.py:13
BEAM_LENGTH = 5                                # ACTIONS['FIRE'] harvest.py:11, cleanup.py:11-12
DEFAULT_VIEW_SIZE = 7                          # harvest.py:15, cleanup.py:22
THRESHOLD_DEPLETION = 0.4                      # cleanup.py:24
THRESHOLD_RESTORATION = 0.0                    # cleanup.py:25
WASTE_SPAWN_PROBABILITY = 0.5                  # cleanup.py:26
APPLE_RESPAWN_PROBABILITY = 0.05               # cleanup.py:27


def colour_lut(colour_map=None):
    """128x3 uint8 RGB table indexed by ASCII code."""
    cm = dict(DEFAULT_COLOURS)
    cm.update(CLEANUP_COLOURS)
    if colour_map:
        cm.update(colour_map)
    lut = np.zeros((128, 3), dtype=np.uint8)
    for ch, rgb in cm.items():
        if len(ch) == 1 and ord(ch) < 128:
            lut[ord(ch)] = rgb
    return lut


def cleanup_probabilities(current_waste, potential_waste_area,
                          threshold_depletion=THRESHOLD_DEPLETION,
                          threshold_restoration=THRESHOLD_RESTORATION,
                          waste_spawn_probability=WASTE_SPAWN_PROBABILITY,
                          apple_respawn_probability=APPLE_RESPAWN_PROBABILITY):
    """(apple_prob, waste_prob) for a given number of 'H' cells -- cleanup.py:156-179, same
    operations in the same order on Python floats (IEEE double)."""
    waste_density = 0
    if potential_waste_area > 0:
        free_area = potential_waste_area - current_waste
        waste_density = 1 - free_area / potential_waste_area
    if waste_density >= threshold_depletion:
        return 0, 0
    if waste_density <= threshold_restoration:
        return apple_respawn_probability, waste_spawn_probability
    spawn_prob = (1 - (waste_density - threshold_restoration)
                  / (threshold_depletion - threshold_restoration)) * apple_respawn_probability
    return spawn_prob, waste_spawn_probability


class EnvConfig(object):
    """Static description of one game: everything `ssd_create` takes."""

    def __init__(self, kind, ascii_map, num_agents, view_size=DEFAULT_VIEW_SIZE,
                 beam_length=BEAM_LENGTH, colour_map=None, harvest_spawn_prob=HARVEST_SPAWN_PROB,
                 cleanup_params=None):
        h, w = validate_map(ascii_map)
        if not 1 <= num_agents <= MAX_AGENTS:
            raise ValueError("num_agents must be in 1..%d" % MAX_AGENTS)
        self.kind, self.height, self.width = kind, h, w
        self.num_agents, self.view_size, self.beam_length = num_agents, view_size, beam_length
        self.ascii_map = list(ascii_map)
        self.base_map = np.array([[ord(ch) for ch in row] for row in ascii_map], dtype=np.uint8)
        if (self.base_map >= 128).any():
            raise ValueError("map characters must be 7-bit ASCII")
        self.colour_lut = colour_lut(colour_map)
        self.harvest_spawn_prob = np.asarray(harvest_spawn_prob, dtype=np.float64)

        def points(chars):
            rr, cc = np.nonzero(np.isin(self.base_map, [ord(c) for c in chars]))
            return np.stack([rr, cc], axis=1).astype(np.int16)  # row-major scan order

        spawn = points('P')                       # map_env.py:96-99
        if kind == KIND_CLEANUP:                  # cleanup.py:51-52 appends every 'P' again
            spawn = np.concatenate([spawn, spawn], axis=0)
            spawn = spawn[np.lexsort((spawn[:, 1], spawn[:, 0]))]
        self.spawn_points = np.ascontiguousarray(spawn)
        self.wall_points = points('@')
        self.apple_points = points('A' if kind == KIND_HARVEST else ('B' if kind == KIND_CLEANUP else ''))
        self.waste_points = points('HR') if kind == KIND_CLEANUP else np.zeros((0, 2), np.int16)
        self.potential_waste_area = int(len(self.waste_points))   # cleanup.py:36-38
        params = cleanup_params or {}
        table = [cleanup_probabilities(hh, self.potential_waste_area, **params)
                 for hh in range(self.potential_waste_area + 1)]
        self.cleanup_apple_prob = np.array([t[0] for t in table], dtype=np.float64)
        self.cleanup_waste_prob = np.array([t[1] for t in table], dtype=np.float64)

    @property
    def view_width(self):
        return 2 * self.view_size + 1

    @property
    def obs_shape(self):
        return (self.num_agents, self.view_width, self.view_width, 3)

    @property
    def max_draws(self):
        """Upper bound of np.random.rand calls in one step (tape row length)."""
        return int(len(self.apple_points) + len(self.waste_points))

    @property
    def num_actions(self):
        return {KIND_HARVEST: 8, KIND_CLEANUP: 9}.get(self.kind, 7)

    def initial_grid(self):
        """world_map right after reset_map()+custom_reset() (map_env.py:560-564,
        harvest.py:57-60, cleanup.py:84-92), ASCII uint8[H, W]."""
        g = np.full(self.base_map.shape, ord(' '), dtype=np.uint8)
        keep = '@' + ('A' if self.kind == KIND_HARVEST else ('HRS' if self.kind == KIND_CLEANUP else ''))
        for ch in keep:
            g[self.base_map == ord(ch)] = ord(ch)
        return g


def make_config(name, num_agents=5, view_size=7, ascii_map=None):
    """EnvConfig of 'harvest' / 'cleanup' with the reference's defaults (harvest.py:20, cleanup.py:32)."""
    name = name.lower()
    if name == "harvest":
        return EnvConfig(KIND_HARVEST, ascii_map or HARVEST_MAP, num_agents, view_size=view_size)
    if name == "cleanup":
        return EnvConfig(KIND_CLEANUP, ascii_map or CLEANUP_MAP, num_agents, view_size=view_size)
    raise ValueError("unknown game %r" % name)
