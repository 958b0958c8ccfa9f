"""CPU: the Python/numpy restatement of the reference loop (oracle/pyloop.py, the timed CPU baseline) against the C oracle
(oracle/hexref.c, pinned to the reference's golden vectors) on the same per-game Philox streams. Bit-exact."""
import numpy as np
import pytest

from oracle import hexref, pyloop
from oracle.philox import GameStream


@pytest.mark.parametrize("N", [3, 5, 11])
@pytest.mark.parametrize("agent_mode", [0, 1, 2])
def test_pyloop_matches_c_oracle(N, agent_mode):
    seed, T = 40 + N, 3 * N * N
    for game in (0, 1, 977):
        ref = hexref.RefBatch(hexref.KIND_SELFPLAY_B, N, 1, seed=seed, game_offset=game, agent_mode=agent_mode)
        robs, _ = ref.reset()
        rng = GameStream(seed, game)
        env = pyloop.SelfPlay(N, None if agent_mode == 2 else agent_mode, rng)
        obs = env.reset()
        assert np.array_equal(obs, robs[0])
        for t in range(T):
            mask = env.legal()
            a = pyloop.random_free_cell(obs, rng)
            obs, r, over = env.step(a)
            o = ref.step(auto_reset=True, want_term=True)
            assert int(o["actions"][0]) == int(a), (game, t)
            assert float(o["reward"][0]) == float(r) and bool(o["done"][0]) == bool(over), (game, t)
            if over:
                assert np.array_equal(o["term_obs"][0], obs), (game, t)
                obs = env.reset()
            assert np.array_equal(o["obs"][0], obs), (game, t)
            assert np.array_equal(o["mask"][0].astype(bool), env.legal()), (game, t)
            e = ref.export()
            assert np.array_equal(e["regions"][0], env.sim.planes), (game, t)
            assert np.array_equal(e["region_counter"][0], env.sim.next_label), (game, t)
            assert int(e["draws"][0]) == rng.idx, (game, t)
        del mask


def test_rate_runs():
    v, sample = pyloop.rate(5, 0.2, 1)
    assert v > 0 and "env steps" in sample
