"""Per-source-line hot spots of one profiled kernel launch.

    python tools/ncu_source_hotspots.py gpurun_out/prof.ncu-rep hex_gym_env_b200/libhexb.so \
        _Z16hexb_step_kernelILi11ELi1EEvN4hexb6ParamsE [source_dir] [top_n]

Joins `ncu -i <rep> --page source --csv` (per-SASS-instruction executed-instruction counts and warp-stall samples; the report
must come from `ncu --set full --import-source on`) with the line table `nvdisasm -g` prints for the same kernel of the same
library build (-lineinfo), and aggregates by source line and by enclosing function. source_dir holds the .cu/.cuh files the
library was built from (default hex_gym_env_b200/csrc).

Since the step kernels are compiled one object per board size from the SAME source file (hexb_step_inst.cu), `cuobjdump -xelf all`
on libhexb.so writes every per-size cubin under one name; pass the object of the size profiled instead of the library
(hex_gym_env_b200/build/step_11.o: same build, same code)."""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

rep, so, kernel = sys.argv[1:4]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
srcdir = sys.argv[4] if len(sys.argv) > 4 else os.path.join(root, "hex_gym_env_b200", "csrc")
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40

with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    dis = None
    for cubin in sorted(glob.glob(os.path.join(tmp, "*.cubin"))):   # one cubin per object file: find the one with the kernel
        syms = subprocess.run(["cuobjdump", "-elf", cubin], stdout=subprocess.PIPE, text=True).stdout
        if kernel in syms:
            dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], stdout=subprocess.PIPE, text=True, check=True).stdout.split("\n")
            if any(l.startswith(kernel + ":") for l in dis):
                break
            dis = None
    assert dis is not None, "kernel %s not found in %s" % (kernel, so)
start = next(i for i, l in enumerate(dis) if l.startswith(kernel + ":"))
loc, seq = None, []
for l in dis[start + 1:]:
    if l.startswith("//--------------------- "):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        loc = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((int(m.group(1), 16), loc))

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# the page holds one section per profiled launch ("Kernel Name" row, header row, one row per SASS instruction): sum the sections
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
hdr = rows[starts[0] + 1]
sections = [[r for r in rows[a + 2:b] if len(r) == len(hdr) and r[0].startswith("0x")] for a, b in zip(starts, starts[1:])]
sections = [sec for sec in sections if len(sec) == len(sections[0])]
data = [list(r) for r in sections[0]]
for col in (hdr.index("Instructions Executed"), hdr.index("# Samples")):
    for k in range(len(data)):
        data[k][col] = str(sum(int(sec[k][col]) for sec in sections))
print("%d profiled launch(es) summed" % len(sections))
ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
assert len(seq) == len(data), "the library is not the build that was profiled (%d vs %d SASS instructions)" % (len(seq), len(data))
base = int(data[0][0], 16)
byline = collections.defaultdict(lambda: [0, 0, 0])
tot_i = tot_s = 0
for (off, lc), r in zip(seq, data):
    assert int(r[0], 16) - base == off
    byline[lc][0] += int(r[ia])
    byline[lc][1] += int(r[isamp])
    byline[lc][2] += 1
    tot_i += int(r[ia])
    tot_s += int(r[isamp])
src = {}
for f in os.listdir(srcdir):
    if f.endswith((".cu", ".cuh", ".h")):
        src[f] = open(os.path.join(srcdir, f)).read().split("\n")


def func_of(f, ln):
    if f not in src:
        return f
    L = src[f]
    for i in range(ln - 1, -1, -1):
        if re.match(r"^(HEXB_HD|__device__|__global__|static)", L[i]) and "(" in L[i]:
            m = re.search(r"(\w+)\s*\(", re.sub(r"__launch_bounds__\([^)]*\)+", "", L[i]))
            return m.group(1) if m else L[i]
    return "?"


print("%s: %d SASS instructions, %d warp instructions executed, %d stall samples" % (kernel, len(seq), tot_i, tot_s))
print("\n-- by source line (share of stall samples, share of executed warp instructions, SASS instructions on the line)")
for (f, ln), (ie, s, cnt) in sorted(byline.items(), key=lambda kv: -kv[1][1])[:top]:
    text = src[f][ln - 1].strip()[:110] if f in src and ln <= len(src[f]) else ""
    print("%6.2f%% samp %6.2f%% instr  n=%3d  %s:%d  %s" % (100.0 * s / tot_s, 100.0 * ie / tot_i, cnt, f, ln, text))
fn = collections.defaultdict(lambda: [0, 0])
for (f, ln), (ie, s, cnt) in byline.items():
    k = func_of(f, ln)
    fn[k][0] += ie
    fn[k][1] += s
print("\n-- by enclosing function")
for k, (ie, s) in sorted(fn.items(), key=lambda kv: -kv[1][1]):
    if s or ie:
        print("%6.2f%% samp %6.2f%% instr  %s" % (100.0 * s / tot_s, 100.0 * ie / tot_i, k))
