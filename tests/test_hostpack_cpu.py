"""Host half of the packed transport (csrc/hexb_hostpack.cpp) without a GPU: every SIMD level the CPU offers, several pool sizes,
ragged word ranges, aligned and unaligned destinations, against a per-cell loop (tests/hostpack_check.cpp)."""
import os
import subprocess
import tempfile

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def binary():
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "hostpack_check")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(HERE, "hostpack_check.cpp")])
        yield exe


@pytest.mark.parametrize("simd", ["0", "2", "512"])
@pytest.mark.parametrize("threads", ["1", "3", "8"])
def test_expand_matches_definition(binary, simd, threads):
    env = dict(os.environ, HEXB_HOST_SIMD=simd, HEXB_HOST_THREADS=threads, HEXB_HOST_SPIN_US="50")
    out = subprocess.run([binary], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert out.returncode == 0, out.stdout
    assert "hostpack ok" in out.stdout
