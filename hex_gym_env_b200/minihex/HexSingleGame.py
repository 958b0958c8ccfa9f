"""Variant B (board -1/+1/0 in the mover's perspective, one ply per step) - same names as minihex/HexSingleGame.py."""
from ..minihex_compat import HexEnvB as HexEnv  # noqa: F401
from ..minihex_compat import HexGameB as HexGame  # noqa: F401
from ..minihex_compat import player_b as player  # noqa: F401
