"""CPU: pins the property checker of tests/properties.py (used at BASELINE's full sizes on the GPU, where the oracle cannot
follow) against the oracle: every property holds on oracle rollouts, and deliberately corrupted outputs are caught."""
import numpy as np
import pytest
import torch

import properties
from oracle import hexref


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


@pytest.mark.parametrize("N,agent_mode", [(3, 2), (5, 0), (7, 1), (11, 2)])
def test_properties_hold_on_oracle_rollouts(N, agent_mode):
    G, T = 300, 2 * N * N // 3 + 5
    ref = hexref.RefBatch(hexref.KIND_SELFPLAY_B, N, G, seed=N, agent_mode=agent_mode)
    ref.reset()
    agent = _t(ref.export()["agent"])
    n_done = 0
    for t in range(T):
        r = ref.step(want_term=True)
        properties.check_selfplay_step(_t(r["obs"]), _t(r["mask"]), _t(r["reward"]), _t(r["done"]), agent, _t(r["term_obs"]))
        n_done += int(r["done"].sum())
        if t % 5 == 0:
            e = ref.export()
            properties.check_exported_labels(_t(e["board"]), _t(e["regions"]), 0)
    assert n_done == int(ref.stats()[0]) and n_done > G


def test_properties_hold_on_variant_A_labels():
    ref = hexref.RefBatch(hexref.KIND_ENV_A, 7, 200, seed=1)
    ref.reset()
    for t in range(30):
        ref.step()
        e = ref.export()
        properties.check_exported_labels(_t(e["board"]), _t(e["regions"]), 2)


def test_checker_catches_corruption():
    N, G = 5, 200
    ref = hexref.RefBatch(hexref.KIND_SELFPLAY_B, N, G, seed=2, agent_mode=2)
    ref.reset()
    agent = _t(ref.export()["agent"])
    for _ in range(6):
        r = ref.step(want_term=True)
    good = [_t(r[k]).clone() for k in ("obs", "mask", "reward", "done")] + [agent, _t(r["term_obs"]).clone()]
    properties.check_selfplay_step(*good)
    d = np.flatnonzero(r["done"])
    live = np.flatnonzero(r["done"] == 0)
    assert len(d) and len(live)

    def broken(i, fn):
        args = [a.clone() for a in good]
        fn(args[i])
        with pytest.raises(AssertionError):
            properties.check_selfplay_step(*args)

    broken(1, lambda m: m.__setitem__((0, 0), 1 - m[0, 0]))                      # mask bit flipped
    broken(2, lambda rw: rw.__setitem__(int(live[0]), 1.0))                       # reward without done
    broken(3, lambda dn: dn.__setitem__(int(d[0]), 0))                            # missed win
    broken(5, lambda tm: tm.__setitem__(int(d[0]), torch.zeros(N, N, dtype=torch.int8)))  # terminal board without a winner

    def extra_stone(o):
        g = int(live[0])
        y, x = np.argwhere(r["obs"][g] == 0)[0]
        o[g, y, x] = -1
    broken(0, extra_stone)                                                        # stone counts (and the mask) disagree

    e = ref.export()
    reg = _t(e["regions"]).clone()
    g, pl, y, x = np.argwhere(e["regions"][:, :, 1:-1, 1:-1] > 2)[0]
    reg[g, pl, y + 1, x + 1] = 99
    ok = True
    try:
        properties.check_exported_labels(_t(e["board"]), reg, 0)
    except AssertionError:
        ok = False
    # a lone stone may legitimately carry any label; a stone inside a group may not
    nb = e["regions"][g, pl, y:y + 3, x:x + 3]
    if (nb != 0).sum() > 1:
        assert not ok


def test_connects_matches_bruteforce():
    rs = np.random.RandomState(0)
    N, G = 6, 400
    stones = rs.rand(G, N, N) < 0.55

    def bfs(s, axis):
        if axis == 1:
            s = s.T
        seen = {(0, x) for x in range(N) if s[0, x]}
        todo = list(seen)
        while todo:
            y, x = todo.pop()
            for dy, dx in ((-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0)):
                v = (y + dy, x + dx)
                if 0 <= v[0] < N and 0 <= v[1] < N and s[v] and v not in seen:
                    seen.add(v)
                    todo.append(v)
        return any(y == N - 1 for y, _ in seen)

    for axis in (0, 1):
        got = properties.connects(_t(stones), axis).numpy()
        want = np.array([bfs(stones[g], axis) for g in range(G)])
        assert np.array_equal(got, want) and want.any() and not want.all()


def test_checksum_is_order_sensitive():
    a = torch.arange(100, dtype=torch.uint8)
    b = a.clone()
    b[[3, 4]] = b[[4, 3]]
    assert properties.checksum(a) == properties.checksum(a.clone()) != properties.checksum(b)
