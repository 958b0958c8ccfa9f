"""CPU: the host side of the opponent pool (hex_gym_env_b200/opponents.py) - the reference's bookkeeping
(minihex/SelfplayWrapper.py:56-67,106-144), the callback's replacement rule (minihex/EvaluationCallback.py:35-50) and the grouped
dispatch of OpponentPolicy.choose_action by the per-game pool entry. The per-episode choice itself is device code, pinned by
tests/golden/oppmodel_*.npz and evalpool_*.npz (tests/test_emu_parity.py here, tests/test_gpu_parity.py on the GPU)."""
import math

import numpy as np
import pytest
import torch

from hex_gym_env_b200.opponents import OpponentPool


class Tagged(object):
    """A batched policy that answers with its own tag for every game it is shown, and counts its calls."""

    def __init__(self, tag):
        self.tag, self.calls, self.seen = tag, 0, 0

    def __call__(self, obs, mask):
        self.calls += 1
        self.seen += obs.shape[0]
        return torch.full((obs.shape[0],), self.tag, dtype=torch.int32)


def _inputs(G, K, seed):
    g = torch.Generator().manual_seed(seed)
    opp_index = torch.randint(-1, K, (G,), generator=g, dtype=torch.int32)
    to_move = torch.randint(0, 3, (G,), generator=g).to(torch.uint8)
    return torch.zeros(G, 3, 3, dtype=torch.int8), torch.ones(G, 9, dtype=torch.uint8), to_move, opp_index


@pytest.mark.parametrize("dense", [True, False])
def test_every_waiting_game_is_answered_by_its_own_entry(dense):
    K, G = 5, 257
    base = Tagged(100)
    pool = OpponentPool(base, buffer_size=K, dense=dense)
    models = [Tagged(200 + k) for k in range(K)]
    for k in (0, 2, 3):
        pool.set_opponent_model(k, models[k], score=0.0)      # entries 1 and 4 still hold the base model, which is also the best
    obs, mask, to_move, opp_index = _inputs(G, K, 1)
    got = pool(obs, mask, to_move, opp_index)
    want = torch.zeros(G, dtype=torch.int32)
    for g in range(G):
        if to_move[g] == 1:
            k = int(opp_index[g])
            want[g] = 100 if k in (-1, 1, 4) else 200 + k
    assert torch.equal(got, want)
    assert got.dtype == torch.int32
    # one call per DISTINCT model: the base model serves three entries with one call
    assert base.calls == 1 and all(models[k].calls == 1 for k in (0, 2, 3)) and models[1].calls == 0
    if not dense:   # the gather form shows every model its own games only
        assert base.seen == int(((to_move == 1) & ((opp_index == -1) | (opp_index == 1) | (opp_index == 4))).sum())


def test_dense_and_gather_agree():
    K, G = 7, 1000
    a, b = OpponentPool(Tagged(1), buffer_size=K, dense=True), OpponentPool(Tagged(1), buffer_size=K, dense=False)
    for k in range(K):
        a.set_opponent_model(k, Tagged(10 + k), 0.0)
        b.set_opponent_model(k, Tagged(10 + k), 0.0)
    x = _inputs(G, K, 2)
    assert torch.equal(a(*x), b(*x))


def test_bookkeeping_follows_the_reference():
    base = Tagged(0)
    pool = OpponentPool(base, buffer_size=4, scores=[0.5, 0.1, 0.3, 0.1])
    assert pool.best_model is base and pool.best_score == 0.5 and pool.best_mean_reward == -np.inf and pool.eval_state is False
    assert len(pool.get_opponent_models()) == 4 and list(pool.get_scores()) == [0.5, 0.1, 0.3, 0.1]
    m = Tagged(1)
    pool.set_opponent_model(1, m, 0.4)                       # better than the entry it replaces, not better than the best
    assert pool.opponent_models[1] is m and pool.opponent_scores[1] == 0.4 and pool.best_model is base
    m2 = Tagged(2)
    pool.set_opponent_model(3, m2, 0.9)
    assert pool.best_model is m2 and pool.best_score == 0.9
    pool.set_eval(True)
    assert pool.eval_state is True
    pool.set_eval(False)
    pool.append_opponent_model(Tagged(3), best_model=True, mean_reward=0.25)
    assert len(pool.opponent_models) == 5 and pool.get_best_mean_reward() == 0.25
    with pytest.raises(AssertionError):                      # set_eval's own assertion: models and scores out of step (:120)
        pool.set_eval(True)
    assert pool.save_best_model() is None                    # an entry without save / save_model


def test_replacement_rule_of_the_callback():
    class Pick(object):
        def __init__(self):
            self.offered = None

        def choice(self, xs):
            self.offered = list(xs)
            return xs[-1]

    pool = OpponentPool(Tagged(0), buffer_size=4, scores=[0.5, 0.1, 0.3, 0.1])
    learner, rng = Tagged(9), Pick()
    score, idx = pool.consider(learner, 0.8, rng=rng)
    assert score == pytest.approx(0.8 * math.exp(0.25 - 1.0)) and rng.offered == [1, 3] and idx == 3
    assert pool.opponent_models[3] is learner and pool.opponent_scores[3] == pytest.approx(score)
    assert pool.best_model is not learner                    # 0.378 does not beat the best score 0.5
    assert pool.consider(Tagged(8), -0.2, rng=rng) == (pytest.approx(-0.2 * math.exp(float(np.mean(pool.opponent_scores)) - 1.0)), None)
    score, idx = pool.consider(Tagged(7), 0.05, rng=rng)     # positive, but no better than the worst entry: nothing happens
    assert idx is None and score < 0.1
    slot = Tagged(5)
    score, idx = pool.consider(Tagged(6), 0.9, rng=rng, place=lambda i: slot)   # the caller decides which object fills the slot
    assert idx is not None and pool.opponent_models[idx] is slot


def test_binding_checks():
    class FakeCfg(object):
        pool_size, eval_state = 3, 0

    class FakeBatch(object):
        manual_opponent, cfg = True, FakeCfg()

        def __init__(self):
            self.calls = []

        def set_eval(self, flag):
            self.calls.append(flag)
            self.cfg.eval_state = int(flag)

    with pytest.raises(ValueError, match="pool_size"):
        OpponentPool(Tagged(0), buffer_size=4, batch=FakeBatch())
    fb = FakeBatch()
    pool = OpponentPool(Tagged(0), buffer_size=3, batch=fb)
    pool.set_eval(True)
    pool.set_eval(False)
    assert fb.calls == [True, False]
    with pytest.raises(ValueError, match="fixed length"):
        pool.append_opponent_model(Tagged(1))
    fb.manual_opponent = False
    with pytest.raises(ValueError, match="manual_opponent"):
        OpponentPool(Tagged(0), buffer_size=3, batch=fb)


def _mlp(sizes, seed):
    import torch.nn as nn
    torch.manual_seed(seed)
    layers = []
    for i, (a, b) in enumerate(zip(sizes[:-1], sizes[1:])):
        layers.append(nn.Linear(a, b))
        if i + 2 < len(sizes):
            layers.append(nn.Tanh())
    return nn.Sequential(*layers)


def test_stacked_mlp_opponents_equal_the_networks_one_by_one():
    """One batched matmul per layer over all slots + a gather == each game's own network applied on its own."""
    import torch.nn as nn
    from hex_gym_env_b200.opponents import StackedMlpOpponents
    N, K, G = 4, 5, 300
    sizes = (N * N, 64, 64, N * N)
    nets = [_mlp(sizes, 100 + k) for k in range(K + 1)]            # slot 0 = the best model, slot k + 1 = entry k
    stack = StackedMlpOpponents(sizes, K + 1, deterministic=True)
    for sl, net in enumerate(nets):
        stack.load(sl, [m for m in net if isinstance(m, nn.Linear)])
    g = torch.Generator().manual_seed(3)
    obs = torch.randint(-1, 2, (G, N, N), generator=g).to(torch.int8)
    mask = (obs.reshape(G, -1) == 0).to(torch.uint8)
    mask[:, 0] = 1                                                   # never an empty mask
    opp_index = torch.randint(-1, K, (G,), generator=g, dtype=torch.int32)
    to_move = torch.ones(G, dtype=torch.uint8)
    pool = OpponentPool(stack.entry(0), buffer_size=K)
    for k in range(K):
        pool.set_opponent_model(k, stack.entry(k + 1), 0.0)
    assert pool._stack_slots(obs.device)[0] is stack
    got = pool(obs, mask, to_move, opp_index)
    lg = stack.logits(obs, (opp_index + 1).long())
    x = obs.reshape(G, -1).float()
    with torch.no_grad():
        for gi in range(G):
            want = nets[int(opp_index[gi]) + 1](x[gi:gi + 1])[0]
            assert torch.allclose(lg[gi], want, atol=1e-5), gi
            legal = want.masked_fill(mask[gi] == 0, float("-inf"))
            top2 = legal.topk(2).values
            if top2[0] - top2[1] > 1e-4:                             # (a tie within rounding may go either way)
                assert int(got[gi]) == int(legal.argmax()), gi
            assert mask[gi, int(got[gi])] == 1
    # a slot as an entry on its own, and a pool with a foreign entry falls back to the per-model dispatch
    one = stack.entry(2)(obs[:7], mask[:7])
    assert torch.equal(one, stack.actions(obs[:7], mask[:7], torch.full((7,), 2)))
    pool.set_opponent_model(1, Tagged(3), 0.0)
    assert pool._stack_slots(obs.device)[0] is None
    mixed = pool(obs, mask, to_move, opp_index)
    assert bool((mixed[opp_index == 1] == 3).all()) and torch.equal(mixed[opp_index != 1], got[opp_index != 1])
    # replacing an entry = loading into its slot: the very same tensors, new numbers; the pool's best slot follows its best model
    before = stack.W[0].data_ptr()
    stack.load(3, [m for m in _mlp(sizes, 999) if isinstance(m, nn.Linear)])
    assert stack.W[0].data_ptr() == before
    pool.set_opponent_model(1, stack.entry(2), 0.7)                  # new best score -> best_model = slot 2
    _, slots = pool._stack_slots(obs.device)
    assert slots.tolist() == [2, 1, 2, 3, 4, 5]
    with pytest.raises(ValueError):
        stack.load(0, [m for m in _mlp((16, 32, 16), 1) if isinstance(m, nn.Linear)])
