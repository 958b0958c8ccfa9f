"""Build a second copy of libhexb.so with extra nvcc flags for same-box A/B runs (tools/ab_probe.py with HEXB_LIB=<path>).
Usage: python tools/build_variant.py <tag> [-DNAME=VALUE ...]   ->  tools/variants/<tag>/libhexb.so (git-ignored, travels with gpurun)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hex_gym_env_b200 import _native
tag, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "variants", tag)
os.makedirs(out, exist_ok=True)
_native.BUILD_DIR = os.path.join(out, "build")
_native.LIB_PATH = os.path.join(out, "libhexb.so")
_native.NVCC_FLAGS = _native.NVCC_FLAGS + flags
print(_native.build(force=True))
