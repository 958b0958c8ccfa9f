"""CPU: the C-ABI library loads and exports every symbol include/hexb.h declares; argument checking and sizing work without
a GPU (no compute call is made here). The product has no CPU path: creating an env without a device must fail loudly."""
import ctypes
import os
import re

import pytest

from hex_gym_env_b200 import _native
from hex_gym_env_b200._native import HexbConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    return _native.lib()


def test_header_and_library_agree(L):
    hdr = open(os.path.join(ROOT, "include", "hexb.h")).read()
    declared = set(re.findall(r"\b(hexb_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_native.SYMBOLS), (declared ^ set(_native.SYMBOLS))
    raw = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert getattr(raw, name) is not None
    assert L.hexb_version() >> 16 == 1


def test_config_struct_matches_header():
    hdr = open(os.path.join(ROOT, "include", "hexb.h")).read()
    body = hdr[hdr.index("typedef struct hexb_config {"):hdr.index("} hexb_config;")]
    fields = re.findall(r"^\s*(?:u?int(?:32|64)_t)\s+([a-z_]+);", body, flags=re.M)
    assert fields == [f[0] for f in HexbConfig._fields_]


def test_state_bytes_and_errors(L):
    def cfg(**kw):
        base = dict(board_size=11, variant=1, num_games=1000, game_offset=0, seed=0, agent_mode=2, opponent_first=0, auto_reset=1,
                    eval_state=0, raw=0, device=0, manual_opponent=0, pool_size=0)
        base.update(kw)
        return HexbConfig(**base)
    c = cfg()
    n = L.hexb_state_bytes(ctypes.byref(c))
    chunks = (1000 + 127) // 128 * 4                       # games padded to 128, 32 per chunk
    assert n >= chunks * 32 * (121 + 4 * 6) and n % 256 == 0   # labels + record words (W + 2 = 6 at 11x11)
    assert L.hexb_host_workspace_bytes(ctypes.byref(c)) >= 1000 * (4 + 121 + 121 + 4 + 1)
    for bad in (cfg(board_size=2), cfg(board_size=20), cfg(variant=2), cfg(num_games=0), cfg(agent_mode=3),
                cfg(variant=0, agent_mode=1), cfg(game_offset=-1), cfg(raw=1, manual_opponent=1), cfg(pool_size=-1)):
        assert L.hexb_state_bytes(ctypes.byref(bad)) == 0
        out = ctypes.c_void_p()
        assert L.hexb_create(ctypes.byref(bad), None, 0, None, ctypes.byref(out)) == -1      # HEXB_ERR_ARG
    out = ctypes.c_void_p()
    assert L.hexb_create(ctypes.byref(c), None, 0, None, ctypes.byref(out)) == -3            # HEXB_ERR_STATE: no buffer
    assert L.hexb_strerror(0) == b"ok" and b"argument" in L.hexb_strerror(-1) and b"unknown" in L.hexb_strerror(-99)
    for fn, args in (("hexb_step", 10), ("hexb_reset", 6), ("hexb_ply", 4), ("hexb_stats", 3), ("hexb_destroy", 1)):
        assert getattr(L, fn)(*([None] * args)) == -1                                         # null handle -> HEXB_ERR_ARG
    assert L.hexb_set_eval(None, 1, None, None) == -1 and L.hexb_half_step(None, 0, None, None, None, None, None) == -1
    assert L.hexb_version() & 0xffff >= 4                                                     # 1.4: hexb_set_eval


def test_mem_alloc_argument_checks(L):
    """hexb_mem_alloc / hexb_mem_free without a device: bad arguments and the missing GPU are reported, nothing is allocated."""
    import torch
    p, got = ctypes.c_void_p(0x1234), ctypes.c_int32(7)
    assert L.hexb_mem_alloc(0, 1 << 20, 1, None, None) == -1                      # HEXB_ERR_ARG: nowhere to put the pointer
    assert L.hexb_mem_alloc(0, 0, 1, ctypes.byref(p), ctypes.byref(got)) == -1    # zero bytes
    assert L.hexb_mem_free(None) == 0                                             # freeing nothing is fine
    if not torch.cuda.is_available():
        assert L.hexb_mem_alloc(0, 1 << 20, 1, ctypes.byref(p), ctypes.byref(got)) == -4   # HEXB_ERR_NOGPU
        assert p.value is None and got.value == 0
        with pytest.raises(RuntimeError, match="hexb_mem_alloc"):
            _native.DeviceBuffer(1 << 20, 0, True)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from hex_gym_env_b200 import HexBatch
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        HexBatch(5, 10)
    from hex_gym_env_b200.rollout import masked_sample
    with pytest.raises(RuntimeError, match="GPU only"):
        masked_sample(torch.zeros(2, 9), torch.ones(2, 9, dtype=torch.uint8))


def test_missing_library_fails_loudly():
    """Without libhexb.so the package must raise, not fall back: load it from a path that does not exist, in a fresh interpreter."""
    import os
    import subprocess
    import sys
    code = ("import hex_gym_env_b200._native as n\n"
            "try:\n    n.lib()\nexcept RuntimeError as e:\n    assert 'no CPU fallback' in str(e).lower() or 'There is no CPU fallback' in str(e), e\n    print('raised')\n"
            "else:\n    raise SystemExit('loaded something')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, HEXB_LIB="/nonexistent/libhexb.so"),
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert out.returncode == 0 and "raised" in out.stdout, out.stderr


def test_every_entry_point_rejects_a_null_handle(L):
    """Every symbol of include/hexb.h that takes a hexb_env* answers HEXB_ERR_ARG for a null handle (no crash, nothing launched)."""
    no_env = {"hexb_version", "hexb_strerror", "hexb_last_cuda_error", "hexb_state_bytes", "hexb_create", "hexb_host_workspace_bytes",
              "hexb_host_packed_bytes", "hexb_host_threads", "hexb_mem_alloc", "hexb_mem_free", "hexb_masked_sample", "hexb_gae"}
    called = 0
    for name, (_res, args) in _native.SYMBOLS.items():
        if name in no_env:
            continue
        vals = [0.0 if a is ctypes.c_double else (0 if a in (ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t) else None) for a in args]
        assert getattr(L, name)(*vals) == -1, name
        called += 1
    assert called == len(_native.SYMBOLS) - len(no_env) and called >= 24


def test_handle_less_entry_points_check_their_arguments(L):
    """hexb_masked_sample / hexb_gae without a handle: null buffers, empty or oversized shapes are HEXB_ERR_ARG before any launch."""
    p = ctypes.c_void_p(8)
    assert L.hexb_gae(None, None, None, 4, 10, 0.99, 0.95, None, None, 0, None) == -1
    assert L.hexb_gae(p, p, p, 0, 10, 0.99, 0.95, p, None, 0, None) == -1 and L.hexb_gae(p, p, p, 4, 0, 0.99, 0.95, p, None, 0, None) == -1
    assert L.hexb_masked_sample(None, None, None, 10, 9, None, None, None, 0, None) == -1
    for G, C in ((10, 0), (10, 400), (0, 9)):
        assert L.hexb_masked_sample(p, p, p, G, C, None, None, None, 0, None) == -1, (G, C)
