"""TEST INFRASTRUCTURE ONLY (see oracle/README.md) -- pure-Python Philox4x32-10.

Independent restatement of the counter-based generator the production kernels use
(Salmon et al., "Parallel Random Numbers: As Easy as 1, 2, 3", SC'11; Random123 v1.14
`philox4x32_R(10, ...)`).  It is checked against the Random123 known-answer vectors in
tests/test_philox.py and is what `oracle/ref_harness.py` feeds into the *unmodified*
Python reference when it replaces `np.random.shuffle`, `np.random.rand`,
`random.shuffle` and `np.random.randint` ("Philox-in" replay).

Stream layout (shared by oracle/ssd_oracle.c and csrc/ssd_kernels.cu):

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (global_env_id, t, stream, block)

    stream 0  STREAM_MOVE    move-priority Fisher-Yates words   (map_env.py:422)
    stream 1  STREAM_SPAWN   uniform draws of the spawn pass    (harvest.py:101, cleanup.py:139,150)
    stream 2  STREAM_WASTE   sort keys of the waste-point order (cleanup.py:145)
    stream 3  STREAM_RPOINT  sort keys of the spawn-point order (map_env.py:656)
    stream 4  STREAM_RROT    spawn rotation                     (map_env.py:666)
    stream 5  STREAM_RSPAWN  uniform draws of reset()'s spawn pass (map_env.py:230)
"""

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF

STREAM_MOVE, STREAM_SPAWN, STREAM_WASTE, STREAM_RPOINT, STREAM_RROT, STREAM_RSPAWN = range(6)


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


class Stream:
    """Word / 53-bit-uniform access into one (seed, env, t, stream) Philox stream."""

    def __init__(self, seed, env_id, t, stream):
        self.key = (seed & MASK, (seed >> 32) & MASK)
        self.env_id, self.t, self.stream = env_id & MASK, t & MASK, stream
        self._cache = {}

    def block(self, b):
        if b not in self._cache:
            self._cache[b] = philox4x32_10((self.env_id, self.t, self.stream, b), self.key)
        return self._cache[b]

    def word(self, i):
        return self.block(i >> 2)[i & 3]

    def u53(self, k):
        """k-th 53-bit integer: numpy's legacy double recipe (a>>5, b>>6) on words (2k, 2k+1)."""
        blk = self.block(k >> 1)
        a, b = blk[2 * (k & 1)], blk[2 * (k & 1) + 1]
        return ((a >> 5) << 26) | (b >> 6)

    def uniform(self, k):
        return self.u53(k) / 9007199254740992.0
