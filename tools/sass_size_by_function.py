"""Static code size of one kernel by enclosing source function: SASS instructions per function / per source line (nvdisasm -g line
table of an object built with -lineinfo). The step kernel's instruction-cache footprint matters (ncu: sm__icc_request_hit_rate 84 %,
'no_instructions' 18 % of the stall samples at 1 Mi games), so this shows where the bytes are.
Usage: python tools/sass_size_by_function.py hex_gym_env_b200/build/step_11.o _Z16hexb_step_kernelILi11ELi1ELb0EEvN4hexb6ParamsE [top_lines]"""
import collections, os, re, subprocess, sys
obj, kernel = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
srcdir = os.path.join(root, "hex_gym_env_b200", "csrc")
import glob, tempfile
with tempfile.TemporaryDirectory() as tmp:   # the host object embeds the cubin
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = sorted(glob.glob(os.path.join(tmp, "*.cubin")))[0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], stdout=subprocess.PIPE, text=True, check=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith(kernel + ":"))
loc, byline, ops = None, collections.Counter(), collections.Counter()
for l in dis[start + 1:]:
    if l.startswith("//--------------------- "):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        loc = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        byline[loc] += 1
src = {f: open(os.path.join(srcdir, f)).read().split("\n") for f in os.listdir(srcdir) if f.endswith((".cu", ".cuh", ".h"))}
def func_of(f, ln):
    if f not in src:
        return f
    L = src[f]
    for i in range(ln - 1, -1, -1):
        if re.match(r"^(HEXB_HD|__device__|__global__|static)", L[i]) and "(" in L[i]:
            m = re.search(r"(\w+)\s*\(", re.sub(r"__launch_bounds__\([^)]*\)+", "", L[i]))
            return m.group(1) if m else L[i]
    return "?"
fn = collections.Counter()
for (f, ln), c in byline.items():
    fn[func_of(f, ln)] += c
tot = sum(byline.values())
print("%s: %d SASS instructions = %.1f KB" % (kernel, tot, tot * 16 / 1024.0))
for k, c in fn.most_common():
    print("%6d  %5.1f%%  %s" % (c, 100.0 * c / tot, k))
print("-- top lines")
for (f, ln), c in byline.most_common(top):
    text = src[f][ln - 1].strip()[:100] if f in src and ln <= len(src[f]) else ""
    print("%6d  %s:%d  %s" % (c, f, ln, text))
