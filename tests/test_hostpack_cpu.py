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


def test_pool_under_thread_sanitizer():
    """The same check built with -fsanitize=thread: the pool's lock-free parts (blocks taken from a shared counter, arrival published
    piece by piece, bounded spinning, one join per step) must be free of data races. Skipped where the toolchain has no libtsan."""
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "hostpack_tsan")
        cc = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-pthread", "-fsanitize=thread", "-o", exe,
                             os.path.join(HERE, "hostpack_check.cpp")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if cc.returncode != 0:
            pytest.skip("no ThreadSanitizer in this toolchain: " + cc.stdout[-200:])
        for threads, simd in (("3", "0"), ("8", "0"), ("8", "0"), ("5", "2"), ("8", "512")):   # (the scalar path's table is built lazily)
            env = dict(os.environ, HEXB_HOST_SIMD=simd, HEXB_HOST_THREADS=threads, HEXB_HOST_SPIN_US="50", TSAN_OPTIONS="halt_on_error=1")
            out = subprocess.run([exe], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
            assert out.returncode == 0 and "hostpack ok" in out.stdout and "WARNING: ThreadSanitizer" not in out.stdout, out.stdout[-2000:]
