"""One-off soak on a GPU box: hundreds of seeds of tests/parity.py::api_fuzz (random walks over the step / reset API against the
oracle) through the C ABI. python tools/soak_fuzz.py [first_seed] [count]"""
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import parity
from gpu_adapter import make
bad = 0
first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
count = int(sys.argv[2]) if len(sys.argv) > 2 else 600
for seed in range(first, first + count):
    try:
        parity.api_fuzz(make, seed, T=40)
    except AssertionError as e:
        bad += 1
        print("FAIL", seed, str(e)[:200])
print("soak done, failures:", bad)
