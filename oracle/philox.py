"""TEST INFRASTRUCTURE ONLY - numpy Philox4x32-10 and the CPython double recipe.

The reference draws from CPython's global ``random`` (MT19937), which cannot be
keyed per game. The batched simulator instead defines one counter-based stream
per game: draw ``i`` of game ``g`` is Philox4x32-10 with counter
``(i, g_lo, g_hi, 0)`` and key ``(seed_lo, seed_hi)``; the first two output
words ``a, b`` become a double exactly the way CPython's ``random.random()``
builds one (``genrand_res53``): ``((a >> 5) * 2**26 + (b >> 6)) / 2**53``.

The same function exists three times and is pinned against the Random123
known-answer vectors in ``tests/test_philox.py``:
  * here (numpy, vectorised over games),
  * ``oracle/hexref.c`` (C),
  * ``hex_gym_env_b200/csrc/hexb_core.cuh`` (device).
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10. All inputs broadcastable integer arrays; returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64) & MASK32
    c1 = np.asarray(c1, dtype=np.uint64) & MASK32
    c2 = np.asarray(c2, dtype=np.uint64) & MASK32
    c3 = np.asarray(c3, dtype=np.uint64) & MASK32
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def draw(seed, game, idx):
    """Draw ``idx`` of game ``game`` under ``seed`` as float64 in [0, 1). Vectorised over game/idx."""
    game = np.asarray(game, dtype=np.uint64)
    idx = np.asarray(idx, dtype=np.uint64)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    a, b, _, _ = philox4x32_10(idx & MASK32, game & MASK32, game >> np.uint64(32), 0,
                               seed & 0xFFFFFFFF, seed >> 32)
    hi = (a >> np.uint32(5)).astype(np.float64)
    lo = (b >> np.uint32(6)).astype(np.float64)
    return (hi * 67108864.0 + lo) / 9007199254740992.0


class GameStream(object):
    """Python ``random``-module look-alike for ONE game: ``random() / uniform() / randint()``.

    ``randint(0, 1)`` (the agent-colour draw of ``SelfplayWrapper.py:72-73``) is defined as
    ``int(random() * 2)`` - one draw, like every other call.
    """

    def __init__(self, seed, game, idx=0):
        self.seed, self.game, self.idx = seed, game, idx

    def random(self):
        u = float(draw(self.seed, self.game, self.idx))
        self.idx += 1
        return u

    def uniform(self, a, b):
        return a + (b - a) * self.random()

    def randint(self, a, b):
        return a + int(self.random() * (b - a + 1))


class ListStream(object):
    """``random`` look-alike that replays a fixed list of doubles (for the SURVEY KATs)."""

    def __init__(self, values):
        self.values, self.idx = list(values), 0

    def random(self):
        u = self.values[self.idx]
        self.idx += 1
        return u

    def uniform(self, a, b):
        return a + (b - a) * self.random()

    def randint(self, a, b):
        return a + int(self.random() * (b - a + 1))
