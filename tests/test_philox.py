"""CPU: Philox4x32-10 - the per-game random stream of the batched simulator - against the published Random123 known-answer
vectors (Random123 examples/kat_vectors, lines `philox4x32 10 ...`), in all three implementations: oracle/philox.py (numpy, what
generated the golden fixtures), oracle/hexref.c (the C oracle) and the host build of csrc/hexb_core.cuh (the code the device
runs, through the emulator). Then the draw construction (CPython's random.random() from two 32-bit words) of all three against
each other."""
import ctypes

import numpy as np
import pytest

from oracle import hexref, philox
from emu import emu

# counter[4], key[2] -> output[4]
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_python_oracle(ctr, key, want):
    out = philox.philox4x32_10(*[np.uint64(c) for c in ctr], key[0], key[1])
    assert tuple(int(o) for o in out) == want


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_c_oracle(ctr, key, want):
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    o = (ctypes.c_uint32 * 4)()
    hexref.lib().hexref_philox4x32_10(c, k, o)
    assert tuple(o) == want


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_device_code_host_build(ctr, key, want):
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    o = (ctypes.c_uint32 * 2)()
    emu.lib().emu_philox4x32_10(c, k, o)
    assert tuple(o) == want[:2]          # the simulator uses the first two output words of a block (one double per block)


def test_draws_agree():
    rs = np.random.RandomState(0)
    for _ in range(300):
        seed = int(rs.randint(0, 2 ** 63 - 1))
        game = int(rs.randint(0, 2 ** 62))
        idx = int(rs.randint(0, 2 ** 31))
        a = float(philox.draw(seed, game, idx))
        b = float(hexref.draw(seed, game, idx))
        c = float(emu.lib().emu_draw01(seed, game, idx))
        assert a == b == c and 0.0 <= a < 1.0
    # the construction itself: ((a >> 5) * 2**26 + (b >> 6)) / 2**53 with the block's first two words (CPython's genrand_res53)
    w = philox.philox4x32_10(np.uint64(5), np.uint64(7), np.uint64(0), 0, 9, 0)
    assert float(philox.draw(9, 7, 5)) == ((int(w[0]) >> 5) * 67108864.0 + (int(w[1]) >> 6)) / 9007199254740992.0
