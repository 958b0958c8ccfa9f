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


# ------------------------------------------------------------------------------------------------ the small methods (a2 / a3 / a7, f4)
@pytest.mark.parametrize("name", golden_files("facade_"))
def test_small_game_methods(mh, name):
    """is_valid_move (incl. IndexError beyond the board), action_to_coordinate, coordinate_to_action, get_possible_actions
    (HexGame.py:74-76,113-122 / HexSingleGame.py:77-79,124-133) at positions along a random game, vs the unmodified reference."""
    minihex, compat, A, B, S = mh
    z = np.load(os.path.join(GOLDEN, name))
    N = int(z["N"])
    C = N * N
    for variant in ("A", "B"):
        g = A.HexGame(A.player.BLACK, A.player.EMPTY * np.ones((N, N)), A.player.BLACK) if variant == "A" else B.HexGame(0, np.zeros((N, N)))
        at = {int(t): i for i, t in enumerate(z["at_" + variant])}
        for t, a in enumerate(z["moves_" + variant]):
            if t in at:
                i = at[t]
                assert [bool(g.is_valid_move(k)) for k in range(C)] == [bool(v) for v in z["valid_" + variant][i]], (name, variant, t)
                pa = np.asarray(g.get_possible_actions())
                n = int(z["npossible_" + variant][i])
                assert len(pa) == n and np.array_equal(pa, z["possible_" + variant][i][:n]), (name, variant, t)
            g.make_move(int(a))
            if variant == "B":
                g._ref_flipped = not g._ref_flipped      # what HexEnv.invert_board does to the reference's board after every ply
        for k, raises in zip(z["oob_actions"], z["oob_raises_" + variant]):
            if raises:
                with pytest.raises(IndexError):
                    g.is_valid_move(int(k))
            else:
                g.is_valid_move(int(k))
        assert np.array_equal(np.array([g.action_to_coordinate(k) for k in range(C)]), z["coords_" + variant])
        assert [int(g.coordinate_to_action(tuple(g.action_to_coordinate(k)))) for k in range(C)] == list(z["actions_of_coords_" + variant])


@pytest.mark.parametrize("name", golden_files("oppredict_"))
def test_opponent_predict(mh, name):
    """Variant-A HexEnv(opponent_policy="opponent_predict", opponent_model=..., eps=0.5) (HexGame.py:165-167, 354-359): draw order
    (uniform, then random_policy's draw when below eps), the board and the mask the model is shown, info dict, rewards."""
    from oracle.scripted import ScriptedModelA
    minihex, compat, A, B, S = mh
    z = np.load(os.path.join(GOLDEN, name))
    N, seed, eps, of = int(z["N"]), int(z["seed"]), float(z["eps"]), int(z["opponent_first"])
    T, G = z["actions"].shape
    import random
    try:
        for gi in range(G):
            stream = GameStream(seed, gi)
            compat.random = stream
            log = []
            env = A.HexEnv(opponent_policy="opponent_predict", opponent_model=ScriptedModelA(log), board_size=N, eps=eps,
                           current_player_num=A.player.WHITE if of else A.player.BLACK)
            obs, _ = env.reset()
            del log[:]
            assert np.array_equal(obs, z["obs0"][gi]) and stream.idx == z["draws0"][gi], (name, gi)
            for t in range(T):
                obs, r, done, trunc, info = env.step(int(z["actions"][t, gi]))
                w = (name, gi, t)
                assert r == z["reward"][t, gi] and bool(done) == bool(z["done"][t, gi]) and trunc is False, w
                lmo = -1 if info["last_move_opponent"] is None else int(info["last_move_opponent"])
                assert lmo == z["last_move_opponent"][t, gi] and code(info["winner"]) == z["winner"][t, gi], w
                assert len(log) == z["model_calls"][t, gi], w
                if log:
                    a, mask, board = log[-1]
                    assert a == z["model_action"][t, gi] and np.array_equal(mask, z["model_mask"][t, gi]), w
                    assert np.array_equal(board, z["model_board"][t, gi]), w
                del log[:]
                assert np.array_equal(env.simulator.regions, z["regions"][t, gi]), w
                assert np.array_equal(env.simulator.region_counter, z["counter"][t, gi]), w
                if done:
                    obs, _ = env.reset()
                    del log[:]
                assert np.array_equal(obs, z["obs"][t, gi]) and stream.idx == z["draws"][t, gi], w
    finally:
        compat.random = random


@pytest.mark.parametrize("name", golden_files("presetreset_"))
def test_env_resets_on_preset_boards(mh, name):
    """HexEnv.reset twice on a preset board, and with user-supplied regions=: the first reset rebuilds the planes, later ones
    adopt the cached planes with region_counter = max(plane) + 1 (HexGame.py:207-220 / HexSingleGame.py:211-229)."""
    minihex, compat, A, B, S = mh
    z = np.load(os.path.join(GOLDEN, name))
    N, tc = int(z["N"]), z["board_true"]
    for i in range(min(len(tc), 12)):
        for variant in ("A", "B"):
            reg, ctr = z["regions_" + variant][i], z["counter_" + variant][i]
            if variant == "A":
                mk = lambda regions=None: A.HexEnv(opponent_policy=None, board=tc[i].astype(np.float64), regions=regions, board_size=N)
            else:
                bb = np.where(tc[i] == 0, -1.0, np.where(tc[i] == 1, 1.0, 0.0))
                mk = lambda regions=None: B.HexEnv(board=bb, regions=regions, board_size=N)
            env = mk()
            for k in range(2):
                env.reset()
                assert np.array_equal(env.simulator.regions, reg[k]) and np.array_equal(env.simulator.region_counter, ctr[k]), (name, i, variant, k)
            env2 = mk(np.array(reg[0], dtype=np.float64))
            env2.reset()
            assert np.array_equal(env2.simulator.regions, reg[2]) and np.array_equal(env2.simulator.region_counter, ctr[2]), (name, i, variant)
            mv = int(z["move_" + variant][i])
            if mv >= 0:
                env.simulator.make_move(mv)
                assert np.array_equal(env.simulator.regions, z["moved_regions_" + variant][i]), (name, i, variant)
                assert np.array_equal(env.simulator.region_counter, z["moved_counter_" + variant][i]), (name, i, variant)
