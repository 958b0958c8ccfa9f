"""`from hex_gym_env_b200 import minihex` - the reference package's import surface (minihex/__init__.py:8-18)."""
from ..minihex_compat import random_policy  # noqa: F401

try:  # register 'hex-v0' like the reference does when gymnasium is available
    from gymnasium.envs.registration import register
    register(id="hex-v0", entry_point="hex_gym_env_b200.minihex.HexGame:HexEnv")
except Exception:  # pragma: no cover - gymnasium is not installed in the build image
    pass
