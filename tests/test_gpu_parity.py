"""GPU (-m gpu): the CUDA library (libhexb.so) called through its C ABI, against the golden vectors of the unmodified
reference and against the oracle on the same seeded inputs. Bit-exact on every array (integer / byte work)."""
import numpy as np
import pytest

import parity
from conftest import golden_files
from oracle import hexref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def make():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    from gpu_adapter import make as mk
    return mk


@pytest.mark.parametrize("name", golden_files("game_A"))
def test_golden_raw_games(make, name):
    parity.golden_raw_game(make, name)


@pytest.mark.parametrize("name", golden_files("selfplay_") + golden_files("envA_"))
def test_golden_rollouts(make, name):
    parity.golden_rollout(make, name)


@pytest.mark.parametrize("N", [3, 4, 5, 6, 7, 8, 9, 11, 13, 16, 19])
@pytest.mark.parametrize("agent_mode", [0, 1, 2])
def test_selfplay_vs_oracle(make, N, agent_mode):
    G = 1000 if N <= 11 else 300   # not a multiple of the 128-game tile: exercises the ragged last tile
    T = 2 * N * N // 3 + 10
    parity.versus_oracle(make, hexref.KIND_SELFPLAY_B, N, G, T, seed=N * 10 + agent_mode, game_offset=12345678901 * agent_mode,
                         fused=(agent_mode != 1), agent_mode=agent_mode, check_state_every=3)


@pytest.mark.parametrize("N", [3, 5, 7, 10, 11])
@pytest.mark.parametrize("opponent_first", [False, True])
def test_envA_vs_oracle(make, N, opponent_first):
    parity.versus_oracle(make, hexref.KIND_ENV_A, N, 700, N * N, seed=5 + N, fused=not opponent_first, opponent_first=opponent_first,
                         check_state_every=3)


@pytest.mark.parametrize("kind", [hexref.KIND_SELFPLAY_B, hexref.KIND_ENV_A])
def test_no_auto_reset(make, kind):
    kw = dict(agent_mode=2) if kind == hexref.KIND_SELFPLAY_B else {}
    parity.versus_oracle(make, kind, 4, 500, 30, seed=3, fused=False, auto_reset=False, illegal_rate=0.1, **kw)


def test_eval_state_draws(make):
    parity.versus_oracle(make, hexref.KIND_SELFPLAY_B, 5, 130, 40, seed=9, agent_mode=2, eval_state=True)


def test_config2_subset(make):
    """BASELINE config 2 (7x7 variant-A HexEnv with action masking): a 4,096-game subset against the oracle."""
    parity.versus_oracle(make, hexref.KIND_ENV_A, 7, 4096, 60, seed=2, fused=True, check_state_every=10)


def test_config3_subset(make):
    """BASELINE config 3 (11x11 SelfPlayEnv, random opponent, agent colour random per game): 8,192 games, 130 steps."""
    parity.versus_oracle(make, hexref.KIND_SELFPLAY_B, 11, 8192, 130, seed=0, fused=True, agent_mode=2, check_state_every=13)
