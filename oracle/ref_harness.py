"""TEST INFRASTRUCTURE ONLY - imports the REAL reference (read-only, /root/reference) in the build container.

The reference's hot path needs only numpy, but its modules import ``gymnasium`` and
``pygame`` unconditionally (``minihex/HexGame.py:1-2,7``, ``minihex/interactive/gui.py:3``),
neither of which is installed. Two tiny stand-ins are placed in ``sys.modules`` so the
reference imports UNMODIFIED. Its module-level ``random`` (``SelfplayWrapper.py:5``,
``minihex/__init__.py:5``, ``HexGame.py:5``) is then swapped for a per-game stream
(``oracle.philox.GameStream``) so draws are reproducible and keyed per game.

This file is used by ``oracle/gen_golden.py``, by optional cross-checks that are skipped when the
reference is absent, and by ``oracle/ref_loop.py`` (the CPU baseline ``bench.py`` times: the
unmodified reference from the byte-identical copy ``oracle/make_ref.py`` puts under ``oracle/_ref``).
Nothing here ships.
"""
import os
import sys
import types

REF_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # oracle/make_ref.py: unmodified copy that travels
REFERENCE_ROOT = os.environ.get("HEX_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "minihex")) and os.path.isdir(os.path.join(REF_COPY, "minihex")):
    REFERENCE_ROOT = REF_COPY   # the GPU box has no /root/reference; the byte-identical copy stands in


def use_copy():
    """Import the reference from oracle/_ref even where /root/reference exists (bench.py: the same files in both places)."""
    global REFERENCE_ROOT
    if _mods is not None and REFERENCE_ROOT != REF_COPY:
        raise RuntimeError("the reference is already imported from %s" % REFERENCE_ROOT)
    REFERENCE_ROOT = REF_COPY


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "minihex"))


_mods = None


def _install_stubs():
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")

        class Env(object):
            pass

        class _Space(object):
            def __init__(self, *a, **k):
                self.args, self.kwargs = a, k

        spaces = types.ModuleType("gymnasium.spaces")
        spaces.Box = _Space
        spaces.Discrete = _Space
        envs = types.ModuleType("gymnasium.envs")
        registration = types.ModuleType("gymnasium.envs.registration")
        registration.register = lambda *a, **k: None
        envs.registration = registration
        gym.Env, gym.spaces, gym.envs = Env, spaces, envs
        sys.modules.update({"gymnasium": gym, "gymnasium.spaces": spaces,
                            "gymnasium.envs": envs, "gymnasium.envs.registration": registration})
    if "pygame" not in sys.modules:
        pg = types.ModuleType("pygame")
        pg.Color = lambda *a, **k: None
        sys.modules["pygame"] = pg


def load():
    """Returns (minihex, HexGame module [variant A], HexSingleGame module [variant B], SelfplayWrapper module)."""
    global _mods
    if _mods is None:
        if not available():
            raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
        _install_stubs()
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        import minihex
        import minihex.HexGame as A
        import minihex.HexSingleGame as B
        import minihex.SelfplayWrapper as S
        _mods = (minihex, A, B, S)
    return _mods


def set_rng(stream):
    """Point every ``random`` the reference hot path uses at ``stream``."""
    minihex, A, B, S = load()
    minihex.random = stream          # minihex/__init__.py:11 (random_policy)
    dup = sys.modules.get("minihex.__init__")   # HexGame.py:6 imports the package body a second time under this name
    if dup is not None:
        dup.random = stream
    A.random = stream                # HexGame.py:355
    S.random = stream                # SelfplayWrapper.py:20,73,97,103,159
