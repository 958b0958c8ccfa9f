"""GPU (-m gpu): the single-environment drop-in classes (hex_gym_env_b200.minihex: HexGame / HexEnv / SelfPlayEnv with the
reference's signatures) replayed against the golden vectors of the unmodified reference, driven exactly the way
oracle/gen_golden.py drove the reference (same loop, same per-game random stream in place of the global `random`)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_files
from oracle.philox import GameStream, ListStream

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mh():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    from hex_gym_env_b200 import minihex, minihex_compat
    from hex_gym_env_b200.minihex import HexGame as A, HexSingleGame as B, SelfplayWrapper as S
    return minihex, minihex_compat, A, B, S


def code(w):
    return -1 if w is None else int(w)


@pytest.mark.parametrize("name", golden_files("game_A"))
def test_hexgame_A_traces(mh, name):
    minihex, compat, A, B, S = mh
    z = np.load(os.path.join(GOLDEN, name))
    N = int(z["N"])
    for gi in range(min(2, z["moves"].shape[0])):
        g = A.HexGame(A.player.BLACK, A.player.EMPTY * np.ones((N, N)), A.player.BLACK)
        for t, a in enumerate(z["moves"][gi]):
            assert code(g.make_move(int(a))) == z["ret"][gi, t]
            assert np.array_equal(g.board, z["board"][gi, t])
            assert np.array_equal(g.regions, z["regions"][gi, t])
            assert np.array_equal(g.region_counter, z["counter"][gi, t])
            assert g.current_player_num == z["cur"][gi, t] and int(g.done) == z["done"][gi, t] and code(g.winner) == z["winner"][gi, t]


@pytest.mark.parametrize("name", golden_files("selfplay_") + golden_files("envA_"))
def test_env_rollouts(mh, name):
    minihex, compat, A, B, S = mh
    z = np.load(os.path.join(GOLDEN, name))
    N, seed, fused = int(z["N"]), int(z["seed"]), int(z["fused"])
    T, G = z["actions"].shape
    T = min(T, 60)
    for gi in range(min(G, 3)):
        stream = GameStream(seed, gi)
        compat.random = stream
        try:
            if name.startswith("selfplay"):
                am = int(z["agent_mode"])
                env = S.selfplay_wrapper(B.HexEnv)(board_size=N, agent_player_num=None if am == 2 else am)
                mask_fn, choose = env.legal_actions, (lambda board: S.BaseRandomPolicy().choose_action(board))
            else:
                env = A.HexEnv(opponent_policy=minihex.random_policy, board_size=N,
                               current_player_num=A.player.WHITE if int(z["opponent_first"]) else A.player.BLACK)
                mask_fn, choose = env.get_action_mask, (lambda board: minihex.random_policy(board))
            obs, _ = env.reset()
            assert np.array_equal(obs, z["obs0"][gi]) and np.array_equal(mask_fn(), z["mask0"][gi].astype(bool))
            assert stream.idx == z["draws0"][gi]
            for t in range(T):
                a = int(choose(obs)) if fused else int(z["actions"][t, gi])
                assert a == z["actions"][t, gi]
                obs, r, done, trunc, _ = env.step(a)
                assert r == z["reward"][t, gi] and bool(done) == bool(z["done"][t, gi]) and trunc is False, (name, gi, t)
                if done:
                    assert np.array_equal(obs, z["term_obs"][t, gi]), (name, gi, t)
                    obs, _ = env.reset()
                assert np.array_equal(obs, z["obs"][t, gi]), (name, gi, t)
                assert np.array_equal(mask_fn(), z["mask"][t, gi].astype(bool)), (name, gi, t)
                assert np.array_equal(env.simulator.regions, z["regions"][t, gi]), (name, gi, t)
                assert np.array_equal(env.simulator.region_counter, z["counter"][t, gi]), (name, gi, t)
                assert env.simulator.current_player_num == z["sim_cur"][t, gi]
                assert stream.idx == z["draws"][t, gi]
        finally:
            import random
            compat.random = random


def test_kats(mh):
    """SURVEY.md section 8c KAT-1..4 through the drop-in classes."""
    minihex, compat, A, B, S = mh
    k = np.load(os.path.join(GOLDEN, "kat.npz"))
    g = A.HexGame(A.player.BLACK, A.player.EMPTY * np.ones((3, 3)), A.player.BLACK)
    rets, empties = [], []
    for m in [4, 0, 1, 3, 7]:
        rets.append(code(g.make_move(m)))
        empties.append(int(g.empty_fields))
    assert rets == list(k["kat1_ret"]) and empties == list(k["kat1_empty_fields"])
    assert np.array_equal(g.board, k["kat1_board"]) and np.array_equal(g.regions, k["kat1_regions"])
    assert np.array_equal(g.region_counter, k["kat1_counter"])
    assert [code(g.make_move(4)), g.current_player_num] == list(k["kat1_again"])
    import random
    try:
        for name, draws, agent, acts in (("kat2", [0.1, 0.2, 0.0, 0.3, 0.99], 0, [4, 1, 7]), ("kat3", [0.1, 0.2, 0.5, 0.5, 0.5], 1, [0, 0])):
            compat.random = ListStream(draws)
            env = S.selfplay_wrapper(B.HexEnv)(board_size=3, agent_player_num=agent)
            obs, _ = env.reset()
            assert np.array_equal(obs, k[name + "_obs"][0])
            for i, a in enumerate(acts):
                obs, r, d, _, _ = env.step(a)
                assert np.array_equal(obs, k[name + "_obs"][i + 1]) and r == k[name + "_r"][i] and bool(d) == bool(k[name + "_d"][i])
            assert np.array_equal(env.simulator.regions, k[name + "_regions"])
            assert [env.current_player_num, env.simulator.current_player_num, code(env.winner)] == list(k[name + "_envcur_simcur_winner"])
        compat.random = ListStream([0.5, 0.5, 0.5])
        env = A.HexEnv(opponent_policy=minihex.random_policy, board_size=3)
        env.reset()
        o1, r1, d1, _, i1 = env.step(4)
        o1 = np.array(o1)
        o2, r2, d2, _, i2 = env.step(4)
        assert np.array_equal(np.array([o1, np.array(o2)]), k["kat4_obs"])
        assert [r1, r2] == list(k["kat4_r"]) and [d1, d2] == [bool(x) for x in k["kat4_d"]]
        assert int(i1["last_move_opponent"]) == int(k["kat4_last_move_opponent"][0]) and code(i2["winner"]) == int(k["kat4_winner"][0])
    finally:
        compat.random = random


def test_debug_mode_raises(mh):
    minihex, compat, A, B, S = mh
    g = B.HexGame(0, np.zeros((4, 4)), debug=True)
    g.make_move(5)
    with pytest.raises(IndexError):
        g.make_move(5)


def test_preset_board_rebuild(mh):
    """HexGame.__init__ with a preset board: raster-order flood_fill rebuild (HexGame.py:53-61) vs the oracle."""
    from oracle import hexref
    minihex, compat, A, B, S = mh
    rs = np.random.RandomState(4)
    for N in (4, 7, 11):
        for _ in range(4):
            board = rs.choice([0, 1, 2], size=(N, N), p=[0.3, 0.3, 0.4]).astype(np.float64)
            g = A.HexGame(A.player.BLACK, board.copy(), A.player.BLACK)
            ref = hexref.RefBatch(hexref.KIND_GAME_A, N, 1)
            ref.set_board(board[None].astype(np.int8), cur=0)
            e = ref.export()
            assert np.array_equal(g.regions, e["regions"][0]) and np.array_equal(g.region_counter, e["region_counter"][0])
            assert np.array_equal(g.board, board)
