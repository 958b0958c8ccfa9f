"""TEST INFRASTRUCTURE ONLY - puts a byte-for-byte copy of the reference's hot-path modules under oracle/_ref/.

    python -m oracle.make_ref            # build container only (needs /root/reference)

Why: BASELINE.json asks for "the reference's Python HexGame loop timed on the GPU box's own host cores", and the GPU box
only receives /root/repo. oracle/_ref/ is git-ignored (the reference's sources never enter this repository's history) but
travels with the working tree, so `bench.py --impl reference` and the `cpu_baseline` leg can drive the UNMODIFIED
reference there (oracle/ref_loop.py) instead of a restatement. Files copied, unchanged:

    minihex/__init__.py  minihex/HexGame.py  minihex/HexSingleGame.py  minihex/SelfplayWrapper.py  minihex/interactive/*.py

oracle/_ref/MANIFEST.json records the sha256 of every source file and of its copy, so "unmodified" is checkable
(tests/test_ref_copy_cpu.py re-hashes them). gymnasium / pygame are not installed anywhere; oracle/ref_harness.py supplies the
two import stubs at run time - nothing is patched into the copied files.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC_ROOT = os.environ.get("HEX_REFERENCE_SRC", "/root/reference")
FILES = ["minihex/__init__.py", "minihex/HexGame.py", "minihex/HexSingleGame.py", "minihex/SelfplayWrapper.py",
         "minihex/interactive/__init__.py", "minihex/interactive/gui.py", "minihex/interactive/interactive.py",
         "minihex/interactive/play_cli.py"]


def sha256(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def source_available():
    return os.path.isfile(os.path.join(SRC_ROOT, FILES[0]))


def present():
    """True when oracle/_ref holds a complete copy whose files still hash to the manifest."""
    man = os.path.join(DEST, "MANIFEST.json")
    if not os.path.isfile(man):
        return False
    try:
        with open(man) as f:
            m = json.load(f)
        return all(sha256(os.path.join(DEST, rel)) == h for rel, h in m["sha256"].items())
    except Exception:
        return False


def make(force=False):
    """Copy the files (no edits) and write the manifest. Returns DEST. No-op when a valid copy exists and not force."""
    if present() and not force:
        return DEST
    if not source_available():
        raise RuntimeError("reference sources not found under %s" % SRC_ROOT)
    hashes = {}
    for rel in FILES:
        src, dst = os.path.join(SRC_ROOT, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        hashes[rel] = sha256(src)
        assert sha256(dst) == hashes[rel]
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "MBPrdctns/hex_gym_env (unmodified files)", "sha256": hashes}, f, indent=1, sort_keys=True)
    return DEST


if __name__ == "__main__":
    print(make(force="--force" in sys.argv))
