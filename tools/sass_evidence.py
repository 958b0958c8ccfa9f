"""SASS mnemonic counts of one kernel of a per-board-size object (hex_gym_env_b200/build/step_N.o): evidence of the TMA bulk
copies (UBLKCP), their mbarrier (SYNCS), programmatic dependent launch (ACQBULK), no local memory (LDL/STL), barriers.
Usage: python tools/sass_evidence.py N mangled_kernel_name [...]"""
import glob, os, re, subprocess, sys, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = int(sys.argv[1])
obj = os.path.join(root, "hex_gym_env_b200", "build", "step_%d.o" % N)
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
    dis = subprocess.run(["nvdisasm", "-c", cubin], stdout=subprocess.PIPE, text=True, check=True).stdout.split("\n")
    res = subprocess.run(["cuobjdump", "-res-usage", cubin], stdout=subprocess.PIPE, text=True).stdout
KEYS = ["UBLKCP", "SYNCS", "FENCE", "REDUX", "SHFL", "VOTE", "LDS", "STS", "STG", "LDG", "LDC", "LDL", "STL", "IMAD.WIDE", "POPC", "PRMT",
        "LOP3", "DMUL", "I2F.F64", "F2I", "REDG", "ATOMG", "BAR.SYNC", "ACQBULK", "UTMACMDFLUSH", "HMMA|UTCHMMA|UTCQMMA"]
for kernel in sys.argv[2:]:
    start = next(i for i, l in enumerate(dis) if l.startswith(kernel + ":"))
    body = []
    for l in dis[start + 1:]:
        if l.startswith("//--------------------- "):
            break
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            body.append(m.group(2))
    print("# %s (step_%d.o, sm_100a): %d SASS instructions" % (kernel, N, len(body)))
    for k in KEYS:
        print("%s: %d" % (k, sum(1 for b in body if re.search(r"(^|\s)(%s)" % k.replace(".", r"\."), b))))
    for l in res.split("\n"):
        if kernel in l:
            i = res.split("\n").index(l)
            print("# " + res.split("\n")[i + 1].strip())
    print("# bulk copies / mbarrier / PDL:")
    for b in body:
        if re.search(r"UBLKCP|SYNCS|ACQBULK|UTMACMDFLUSH|BAR\.SYNC", b):
            print("    " + b)
    print()
