"""ctypes binding of libhexb.so (the C ABI declared in include/hexb.h).

The library is built in-tree by :func:`build` (``nvcc -gencode arch=compute_100a,code=sm_100a``). There is no CPU
implementation behind this module: if the shared library cannot be built or loaded, importing a simulator class fails
with a RuntimeError.
"""
import ctypes
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libhexb.so")
SOURCES = [os.path.join(_CSRC, f) for f in ("hexb_kernels.cu", "hexb_core.cuh", "hexb_phases.cuh", "hexb_views.cuh")] + \
          [os.path.join(os.path.dirname(_HERE), "include", "hexb.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build(force=False, verbose=False):
    """Compile libhexb.so for sm_100a if it is missing or older than its sources. Returns the library path."""
    def stale():
        return (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in SOURCES)

    if force or stale():
        import fcntl
        with open(LIB_PATH + ".lock", "w") as lock:     # one builder at a time (torchrun starts several ranks at once)
            fcntl.flock(lock, fcntl.LOCK_EX)
            if force or stale():                        # another rank may have built it while we waited
                tmp = "%s.%d.tmp" % (LIB_PATH, os.getpid())
                cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp, SOURCES[0]]
                proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                if proc.returncode != 0:
                    if os.path.exists(tmp):
                        os.remove(tmp)
                    raise RuntimeError("building libhexb.so failed:\n%s\n%s" % (" ".join(cmd), proc.stdout))
                os.replace(tmp, LIB_PATH)               # atomic: a concurrent loader never sees a half-written library
                if verbose:
                    print(proc.stdout)
    return LIB_PATH


class HexbConfig(ctypes.Structure):
    _fields_ = [("board_size", ctypes.c_int32), ("variant", ctypes.c_int32), ("num_games", ctypes.c_int64),
                ("game_offset", ctypes.c_int64), ("seed", ctypes.c_uint64), ("agent_mode", ctypes.c_int32),
                ("opponent_first", ctypes.c_int32), ("auto_reset", ctypes.c_int32), ("eval_state", ctypes.c_int32),
                ("raw", ctypes.c_int32), ("device", ctypes.c_int32), ("manual_opponent", ctypes.c_int32),
                ("pool_size", ctypes.c_int32)]


# every symbol include/hexb.h declares: name -> (restype, argtypes)
_vp, _i32, _sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_size_t
_cfgp = ctypes.POINTER(HexbConfig)
SYMBOLS = {
    "hexb_version": (_i32, []),
    "hexb_strerror": (ctypes.c_char_p, [_i32]),
    "hexb_last_cuda_error": (_i32, []),
    "hexb_state_bytes": (_sz, [_cfgp]),
    "hexb_create": (_i32, [_cfgp, _vp, _sz, _vp, ctypes.POINTER(_vp)]),
    "hexb_destroy": (_i32, [_vp]),
    "hexb_get_config": (_i32, [_vp, _cfgp]),
    "hexb_reset": (_i32, [_vp] * 6),
    "hexb_step": (_i32, [_vp] * 10),
    "hexb_host_workspace_bytes": (_sz, [_cfgp]),
    "hexb_step_host": (_i32, [_vp] * 8),
    "hexb_ply": (_i32, [_vp] * 4),
    "hexb_encode": (_i32, [_vp, _i32, _vp, _vp, _vp]),
    "hexb_sample_actions": (_i32, [_vp, _i32, _vp, _vp, _vp]),
    "hexb_export_state": (_i32, [_vp] * 10),
    "hexb_import_boards": (_i32, [_vp] * 5),
    "hexb_stats": (_i32, [_vp] * 3),
    "hexb_rollout": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hexb_set_opponent_buffers": (_i32, [_vp, _vp, _vp]),
    "hexb_set_info_buffers": (_i32, [_vp, _vp, _vp]),
    "hexb_half_step": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "hexb_masked_sample": (_i32, [_vp, _vp, _vp, ctypes.c_int64, _i32, _vp, _vp, _vp, _i32, _vp]),
}

_LIB = None


def lib():
    """The loaded library. Raises RuntimeError (never falls back) if it cannot be built or loaded."""
    global _LIB
    if _LIB is None:
        try:
            # HEXB_LIB: load this build of the same CUDA library instead (kernel-tuning experiments, e.g. other CTA shapes)
            path = os.environ.get("HEXB_LIB") or build()
            L = ctypes.CDLL(path)
        except (OSError, RuntimeError) as exc:
            raise RuntimeError("hex_gym_env_b200: the CUDA library libhexb.so is required and could not be loaded "
                               "(%s). There is no CPU fallback." % (exc,))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _LIB = L
    return _LIB


class HexbError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        L = lib()
        msg = L.hexb_strerror(rc).decode()
        if rc == -2:
            msg += " (cudaError %d)" % L.hexb_last_cuda_error()
        raise HexbError("libhexb: %s" % msg)
