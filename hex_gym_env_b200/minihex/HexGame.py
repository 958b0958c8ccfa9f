"""Variant A (board 0/1/2, opponent inside step) - same names as minihex/HexGame.py of the reference."""
from ..minihex_compat import HexEnvA as HexEnv  # noqa: F401
from ..minihex_compat import HexGameA as HexGame  # noqa: F401
from ..minihex_compat import player, random_policy  # noqa: F401
