"""The C ABI used from plain C (examples/c_abi_demo.c: no Python, no torch on the call path). CPU: it compiles and links against
libhexb.so. GPU: it runs and its episode statistics equal the oracle's for the same seed (the simulation is deterministic)."""
import os
import subprocess

import numpy as np
import pytest

from hex_gym_env_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "c_abi_demo")


def build_demo():
    _native.build()
    cmd = ["gcc", "-O2", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", os.path.join(ROOT, "examples", "c_abi_demo.c"),
           "-o", EXE, "-L", os.path.join(ROOT, "hex_gym_env_b200"), "-lhexb", "-L", "/usr/local/cuda/lib64", "-lcudart",
           "-Wl,-rpath," + os.path.join(ROOT, "hex_gym_env_b200")]
    subprocess.check_call(cmd)
    return EXE


def test_demo_compiles_and_links():
    exe = build_demo()
    assert os.access(exe, os.X_OK)


@pytest.mark.gpu
def test_demo_matches_oracle():
    from oracle import hexref
    exe = build_demo()
    G, T, seed = 4099, 150, 12
    out = subprocess.run([exe, str(G), str(T), str(seed)], stdout=subprocess.PIPE, text=True, check=True).stdout
    got = [int(x) for x in out.splitlines()[0].split()[1:]]
    ref = hexref.RefBatch(hexref.KIND_SELFPLAY_B, 11, G, seed=seed, agent_mode=2)
    ref.reset()
    for _ in range(T):
        ref.step_fast()
    assert got == ref.stats().tolist(), (got, ref.stats().tolist())
    assert got[6] == G * T and got[0] > 0
