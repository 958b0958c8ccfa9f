"""CPU: the C oracle (oracle/hexref.c) against the golden vectors produced by the unmodified reference
(tests/golden/*.npz, generator: oracle/gen_golden.py). Bit-exact on every array."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_files
from oracle import hexref


def load(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.mark.parametrize("name", golden_files("game_"))
def test_raw_game_traces(name):
    z = load(name)
    N = int(z["N"])
    variant = name.split("_")[1]
    moves = z["moves"]
    n_games, T = moves.shape
    b = hexref.RefBatch(hexref.KIND_GAME_A if variant == "A" else hexref.KIND_GAME_B, N, n_games)
    for t in range(T):
        ret = b.ply(moves[:, t])
        e = b.export()
        assert np.array_equal(ret, z["ret"][:, t]), (name, t)
        assert np.array_equal(e["board"], z["board"][:, t].astype(np.float64)), (name, t)
        assert np.array_equal(e["regions"], z["regions"][:, t].astype(np.float64)), (name, t)
        assert np.array_equal(e["region_counter"], z["counter"][:, t].astype(np.float64)), (name, t)
        assert np.array_equal(e["cur"], z["cur"][:, t]), (name, t)
        assert np.array_equal(e["done"], z["done"][:, t]), (name, t)
        assert np.array_equal(e["winner"], z["winner"][:, t]), (name, t)


def _check_rollout(name, kind):
    z = load(name)
    N, seed, fused = int(z["N"]), int(z["seed"]), int(z["fused"])
    T, G = z["actions"].shape
    kw = {}
    if kind == hexref.KIND_SELFPLAY_B:
        kw["agent_mode"] = int(z["agent_mode"])
    else:
        kw["opponent_first"] = bool(int(z["opponent_first"]))
    b = hexref.RefBatch(kind, N, G, seed=seed, **kw)
    obs0, mask0 = b.reset()
    assert np.array_equal(obs0, z["obs0"]), name
    assert np.array_equal(mask0, z["mask0"]), name
    e = b.export()
    assert np.array_equal(e["agent"], z["agent"]), name
    assert np.array_equal(e["draws"], z["draws0"]), name
    for t in range(T):
        o = b.step(None if fused else z["actions"][t], auto_reset=True, want_term=True)
        assert np.array_equal(o["actions"], z["actions"][t]), (name, t)
        assert np.array_equal(o["reward"], z["reward"][t]), (name, t)
        assert np.array_equal(o["done"], z["done"][t]), (name, t)
        assert np.array_equal(o["obs"], z["obs"][t]), (name, t)
        assert np.array_equal(o["mask"], z["mask"][t]), (name, t)
        d = z["done"][t].astype(bool)
        assert np.array_equal(o["term_obs"][d], z["term_obs"][t][d]), (name, t)
        e = b.export()
        assert np.array_equal(e["regions"], z["regions"][t].astype(np.float64)), (name, t)
        assert np.array_equal(e["region_counter"], z["counter"][t].astype(np.float64)), (name, t)
        assert np.array_equal(e["cur"], z["sim_cur"][t]), (name, t)
        assert np.array_equal(e["draws"], z["draws"][t]), (name, t)


@pytest.mark.parametrize("name", golden_files("selfplay_"))
def test_selfplay_rollouts(name):
    _check_rollout(name, hexref.KIND_SELFPLAY_B)


@pytest.mark.parametrize("name", golden_files("envA_"))
def test_envA_rollouts(name):
    _check_rollout(name, hexref.KIND_ENV_A)


@pytest.mark.parametrize("name", golden_files("oppmodel_") + golden_files("evalpool_"))
def test_scripted_opponent_rollouts(name):
    """The caller-driven opponent path of the oracle against the reference run with OpponentPolicy opponents."""
    import parity
    parity.golden_oppmodel(lambda kind, N, G, **kw: hexref.RefBatch(kind, N, G, **kw), name)


@pytest.mark.parametrize("name", golden_files("preset_"))
def test_preset_boards(name):
    import parity

    def make_raw(kind, N, G):
        z = load(name)
        tc = z["board_true"]
        b = hexref.RefBatch(kind, N, G)
        b.set_board(tc if kind == hexref.KIND_GAME_A else np.where(tc == 0, -1, np.where(tc == 1, 1, 0)).astype(np.int8), cur=0)
        return b
    parity.golden_preset(make_raw, name)


def test_kats():
    """SURVEY.md section 8c KAT-1..4 (values stored from the reference in kat.npz)."""
    k = load("kat.npz")
    # KAT-1: raw variant-A game, moves [4,0,1,3,7]
    b = hexref.RefBatch(hexref.KIND_GAME_A, 3, 1)
    rets = [int(b.ply(np.array([m]))[0]) for m in [4, 0, 1, 3, 7]]
    assert rets == list(k["kat1_ret"]) == [-1, -1, -1, -1, 0]
    e = b.export()
    assert np.array_equal(e["board"][0], k["kat1_board"])
    assert np.array_equal(e["regions"][0], k["kat1_regions"])
    assert np.array_equal(e["region_counter"][0], k["kat1_counter"]) and list(k["kat1_counter"]) == [4, 3]
    assert int(b.ply(np.array([4]))[0]) == 3 and int(b.export()["cur"][0]) == 1 == int(k["kat1_again"][1])
    # KAT-2 / KAT-3: SelfPlayEnv with injected draws (only the draws that reach the board matter)
    for name, agent, open_u, acts, us in (("kat2", 0, 0.0, [4, 1, 7], [0.0, 0.99, 0.0]), ("kat3", 1, 0.5, [0, 0], [0.5, 0.0])):
        b = hexref.RefBatch(hexref.KIND_SELFPLAY_B, 3, 1, agent_mode=agent)
        obs, _ = b.reset(open_u=np.array([open_u]))
        assert np.array_equal(obs[0], k[name + "_obs"][0])
        for i, (a, u) in enumerate(zip(acts, us)):
            o = b.step(np.array([a]), opp_u=np.array([[u, 0.0]]), auto_reset=False)
            assert np.array_equal(o["obs"][0], k[name + "_obs"][i + 1]), (name, i)
            assert o["reward"][0] == k[name + "_r"][i] and bool(o["done"][0]) == bool(k[name + "_d"][i])
        assert np.array_equal(b.export()["regions"][0], k[name + "_regions"])
    # KAT-4: variant-A env, opponent u = 0.5
    b = hexref.RefBatch(hexref.KIND_ENV_A, 3, 1)
    b.reset()
    o1 = b.step(np.array([4]), opp_u=np.array([[0.5, 0.0]]), auto_reset=False)
    o2 = b.step(np.array([4]), opp_u=np.array([[0.5, 0.0]]), auto_reset=False)
    assert np.array_equal(o1["obs"][0], k["kat4_obs"][0]) and np.array_equal(o2["obs"][0], k["kat4_obs"][1])
    assert [o1["reward"][0], o2["reward"][0]] == list(k["kat4_r"]) == [0, -100]
    assert [bool(o1["done"][0]), bool(o2["done"][0])] == [False, True]


@pytest.mark.parametrize("name", golden_files("saturation_"))
def test_label_saturation(name):
    """The oracle against the reference's snapshots of the label-range stress game (oracle/gen_golden.py: saturation_moves)."""
    z = load(name)
    N = int(z["N"])
    b = hexref.RefBatch(hexref.KIND_GAME_A, N, 1)
    snaps = {int(t): i for i, t in enumerate(z["snap_t"])}
    for t, a in enumerate(z["moves"]):
        assert b.ply(np.array([a], np.int32))[0] == z["ret"][t], (name, t)
        if t in snaps:
            e, i = b.export(), snaps[t]
            assert np.array_equal(e["regions"][0], z["regions"][i].astype(np.float64)), (name, t)
            assert np.array_equal(e["region_counter"][0], z["counter"][i].astype(np.float64)), (name, t)
            assert np.array_equal(e["board"][0], z["board"][i].astype(np.float64)), (name, t)
    assert z["counter"].max() >= (105 if N >= 19 else 95)


@pytest.mark.parametrize("name", golden_files("presetreset_"))
def test_preset_resets(name):
    """HexGame.__init__ with and without connected_stones: raster-order rebuild vs adopted planes (different counters)."""
    z = load(name)
    N, tc = int(z["N"]), z["board_true"]
    for variant, kind in (("A", hexref.KIND_GAME_A), ("B", hexref.KIND_GAME_B)):
        boards = tc if variant == "A" else np.where(tc == 0, -1, np.where(tc == 1, 1, 0)).astype(np.int8)
        b = hexref.RefBatch(kind, N, len(tc))
        b.set_board(boards)
        e = b.export()
        assert np.array_equal(e["regions"], z["regions_" + variant][:, 0].astype(np.float64))
        assert np.array_equal(e["region_counter"], z["counter_" + variant][:, 0].astype(np.float64))
        for k in (1, 2):      # second reset (cached planes) and user-supplied regions=
            b.set_board_labels(boards, z["regions_" + variant][:, 0])
            e = b.export()
            assert np.array_equal(e["regions"], z["regions_" + variant][:, k].astype(np.float64))
            assert np.array_equal(e["region_counter"], z["counter_" + variant][:, k].astype(np.float64))
        mv = z["move_" + variant]
        ok = mv >= 0
        b.ply(np.where(ok, mv, 0).astype(np.int32))
        e = b.export()
        assert np.array_equal(e["regions"][ok], z["moved_regions_" + variant][ok].astype(np.float64))
        assert np.array_equal(e["region_counter"][ok], z["moved_counter_" + variant][ok].astype(np.float64))


@pytest.mark.parametrize("name", golden_files("oppredict_"))
def test_opponent_predict_batched(name):
    """Variant A with a caller-driven, eps-mixed opponent (HexEnv.opponent_predict) in the oracle against the reference's run."""
    import parity
    parity.golden_oppredict_batched(lambda kind, N, G, **kw: hexref.RefBatch(kind, N, G, **kw), name)
