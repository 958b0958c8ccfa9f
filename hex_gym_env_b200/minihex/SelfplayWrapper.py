"""Self-play wrapper and policies - same names as minihex/SelfplayWrapper.py of the reference."""
from ..minihex_compat import BaseRandomPolicy, OpponentPolicy, selfplay_wrapper  # noqa: F401
from ..minihex_compat import player_b as player  # noqa: F401
