"""CPU: the C oracle and the device logic (emulator) against the UNMODIFIED reference run live on seeds that are NOT among the
committed fixtures. oracle/gen_golden.py's own drivers execute the reference (from /root/reference where it exists, else from the
byte-identical copy oracle/_ref that oracle/make_ref.py writes and tests/test_ref_copy_cpu.py verifies); the traces go to a
temporary directory in the fixtures' format and through the same checkers as the committed ones. Skipped only where neither
source of the reference exists."""
import os

import numpy as np
import pytest

import parity
from oracle import hexref
from oracle import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason="neither /root/reference nor oracle/_ref is present")


@pytest.fixture(scope="module")
def gg():
    from oracle import gen_golden
    return gen_golden


def _oracle(kind, N, G, **kw):
    return hexref.RefBatch(kind, N, G, **kw)


def _emu(kind, N, G, **kw):
    from test_emu_parity import make
    return make(kind, N, G, **kw)


@pytest.mark.parametrize("N,G,T,seed,agent_mode,fused", [(4, 10, 40, 90001, 2, True), (6, 8, 60, 90002, 0, False), (9, 4, 90, 90003, 1, True),
                                                      (11, 4, 140, 90004, 2, True), (13, 2, 120, 90005, 2, False)])
def test_selfplay_rollouts_on_fresh_seeds(gg, tmp_path, N, G, T, seed, agent_mode, fused):
    from test_oracle_golden import _check_rollout
    o = gg.rollout("B", N, G, T, seed=seed, agent_mode=agent_mode, fused=fused)
    path = str(tmp_path / "selfplay_live.npz")
    np.savez_compressed(path, N=N, seed=seed, agent_mode=agent_mode, fused=int(fused), **o)
    _check_rollout(path, hexref.KIND_SELFPLAY_B)
    parity.golden_rollout(_emu, path)           # the device logic against the same live trace
    assert o["done"].sum() > 0


@pytest.mark.parametrize("N,G,T,seed,of,fused", [(4, 10, 40, 90011, 0, True), (6, 8, 50, 90012, 1, False), (7, 6, 70, 90013, 0, False),
                                              (11, 3, 130, 90014, 1, True)])
def test_envA_rollouts_on_fresh_seeds(gg, tmp_path, N, G, T, seed, of, fused):
    from test_oracle_golden import _check_rollout
    o = gg.rollout("A", N, G, T, seed=seed, agent_mode=0, fused=fused, opponent_first=bool(of))
    path = str(tmp_path / "envA_live.npz")
    np.savez_compressed(path, N=N, seed=seed, opponent_first=of, fused=int(fused), **o)
    _check_rollout(path, hexref.KIND_ENV_A)
    parity.golden_rollout(_emu, path)


@pytest.mark.parametrize("make", [_oracle, _emu], ids=["oracle", "emulator"])
@pytest.mark.parametrize("N,G,T,seed,agent_mode,pool,sched", [(4, 8, 70, 90021, 2, 4, {10: True, 50: False}),
                                                             (5, 6, 60, 90022, 1, 3, None)])
def test_learned_opponents_and_evaluation_cycle_on_fresh_seeds(gg, tmp_path, make, N, G, T, seed, agent_mode, pool, sched):
    o = gg.rollout_scripted_opponent(N, G, T, seed=seed, agent_mode=agent_mode, pool=pool, eval_schedule=sched)
    path = str(tmp_path / "oppmodel_live.npz")
    np.savez_compressed(path, N=N, seed=seed, agent_mode=agent_mode, pool=pool, **o)
    parity.golden_oppmodel(make, path)


@pytest.mark.parametrize("make", [_oracle, _emu], ids=["oracle", "emulator"])
@pytest.mark.parametrize("N,G,T,seed,eps,of", [(4, 10, 30, 90031, 0.3, True), (5, 8, 40, 90032, 0.7, False)])
def test_batched_opponent_predict_on_fresh_seeds(gg, tmp_path, monkeypatch, make, N, G, T, seed, eps, of):
    monkeypatch.setattr(gg, "OUT", str(tmp_path))
    gg.gen_opponent_predict(N, G, T, seed=seed, eps=eps, opponent_first=of)
    parity.golden_oppredict_batched(make, str(tmp_path / ("oppredict_N%d_of%d.npz" % (N, int(of)))))


@pytest.mark.parametrize("variant,N,n,seed", [("A", 4, 6, 90041), ("A", 8, 3, 90042), ("B", 5, 5, 90043), ("B", 10, 2, 90044),
                                              ("A", 19, 1, 90045), ("B", 19, 1, 90046)])
def test_raw_game_traces_on_fresh_seeds(gg, tmp_path, monkeypatch, variant, N, n, seed):
    """Raw HexGame.make_move traces (random moves incl. occupied cells, played on past the win) of the live reference: the oracle
    for both variants, the device logic for variant A (the raw handle of the C ABI is the variant-A game)."""
    from test_oracle_golden import test_raw_game_traces as check_oracle
    monkeypatch.setattr(gg, "OUT", str(tmp_path))
    gg.gen_raw_games(variant, N, n, seed)
    path = str(tmp_path / ("game_%s_N%d.npz" % (variant, N)))
    monkeypatch.setattr("test_oracle_golden.load", lambda name: np.load(path))
    check_oracle("game_%s_N%d.npz" % (variant, N))
    if variant == "A":
        parity.golden_raw_game(_emu, path)
