"""GPU (-m gpu): torch.ops.hexb.* (the C ABI as PyTorch operators) against the ctypes binding and the oracle."""
import numpy as np
import pytest
import torch

from oracle import hexref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    from hex_gym_env_b200 import torch_ops
    return torch_ops.load()


def _bufs(G, N, T=None):
    lead = () if T is None else (T,)
    dev = "cuda"
    return dict(obs=torch.empty(lead + (G, N, N), dtype=torch.int8, device=dev), mask=torch.empty(lead + (G, N * N), dtype=torch.uint8, device=dev),
                reward=torch.empty(lead + (G,), dtype=torch.float32, device=dev), done=torch.empty(lead + (G,), dtype=torch.uint8, device=dev),
                term=torch.zeros(lead + (G, N, N), dtype=torch.int8, device=dev), act=torch.empty(lead + (G,), dtype=torch.int32, device=dev))


def test_step_ops_match_oracle(ops):
    from hex_gym_env_b200 import HexBatch
    N, G, T = 11, 1000, 90
    env = HexBatch(N, G, variant=1, device=0, seed=4, agent_mode=2)
    ref = hexref.RefBatch(hexref.KIND_SELFPLAY_B, N, G, seed=4, agent_mode=2)
    b = _bufs(G, N)
    ops.reset(env.handle, None, None, b["obs"], b["mask"])
    ro, rm = ref.reset()
    assert np.array_equal(b["obs"].cpu().numpy(), ro) and np.array_equal(b["mask"].cpu().numpy(), rm)
    for t in range(T):
        ops.step(env.handle, None, None, b["obs"], b["mask"], b["reward"], b["done"], b["term"], b["act"])
        r = ref.step(want_term=True)
        for k, rk in (("obs", "obs"), ("mask", "mask"), ("reward", "reward"), ("done", "done"), ("act", "actions")):
            assert np.array_equal(b[k].cpu().numpy(), r[rk]), (k, t)
        d = r["done"].astype(bool)
        assert np.array_equal(b["term"].cpu().numpy()[d], r["term_obs"][d])
    st = torch.empty(8, dtype=torch.int64, device="cuda")
    ops.stats(env.handle, st)
    assert np.array_equal(st.cpu().numpy(), ref.stats())


def test_ops_equal_ctypes_binding_with_external_actions(ops):
    from hex_gym_env_b200 import HexBatch
    N, G = 7, 515
    a = HexBatch(N, G, variant=0, device=0, seed=9)
    c = HexBatch(N, G, variant=0, device=0, seed=9)
    b = _bufs(G, N)
    a.reset()
    ops.reset(c.handle, None, None, b["obs"], b["mask"])
    u = torch.rand(G, dtype=torch.float64, device="cuda")
    acts = torch.empty(G, dtype=torch.int32, device="cuda")
    for t in range(40):
        u.uniform_()
        ops.sample_actions(c.handle, 0, u, acts)
        assert torch.equal(acts, a.sample_actions(u))
        o = a.step(acts)
        ops.step(c.handle, acts, None, b["obs"], b["mask"], b["reward"], b["done"], None, None)
        for k in ("obs", "mask", "reward", "done"):
            assert torch.equal(o[k], b[k]), (k, t)
    eo, em = a.encode(1)
    o1, m1 = torch.empty_like(b["obs"]), torch.empty_like(b["mask"])
    ops.encode(c.handle, 1, o1, m1)
    assert torch.equal(eo, o1) and torch.equal(em, m1)


def test_rollout_ply_half_step_ops(ops):
    from hex_gym_env_b200 import HexBatch
    N, G, T = 6, 300, 12
    a = HexBatch(N, G, variant=1, device=0, seed=1, agent_mode=2)
    c = HexBatch(N, G, variant=1, device=0, seed=1, agent_mode=2)
    a.reset(); c.reset()
    b = _bufs(G, N, T)
    ops.rollout(c.handle, T, b["obs"], b["mask"], b["reward"], b["done"], None, b["act"])
    for t in range(T):
        o = a.step(want_actions=True)
        for k, ok in (("obs", "obs"), ("mask", "mask"), ("reward", "reward"), ("done", "done"), ("act", "actions")):
            assert torch.equal(b[k][t], o[ok]), (k, t)
    raw_a, raw_c = HexBatch(5, 64, variant=0, device=0, raw=True), HexBatch(5, 64, variant=0, device=0, raw=True)
    raw_a.reset(); raw_c.reset()
    ret = torch.empty(64, dtype=torch.int8, device="cuda")
    for t in range(12):
        mv = torch.randint(0, 25, (64,), dtype=torch.int32, device="cuda")
        ops.ply(raw_c.handle, mv, ret)
        assert torch.equal(ret, raw_a.ply(mv))
    ha = HexBatch(5, 200, variant=1, device=0, seed=2, agent_mode=2, manual_opponent=True)
    hc = HexBatch(5, 200, variant=1, device=0, seed=2, agent_mode=2, manual_opponent=True)
    ha.reset(); hc.reset()
    rew, dn = torch.empty(200, dtype=torch.float32, device="cuda"), torch.empty(200, dtype=torch.uint8, device="cuda")
    for t in range(20):
        for side in (1, 0):
            h = ha.half_step(side, None) if side == 1 else ha.half_step(0, ha.sample_actions(np.full(200, 0.3)))
            acts = None if side == 1 else hc.sample_actions(np.full(200, 0.3))
            ops.half_step(hc.handle, side, acts, rew, dn, None)
            assert torch.equal(h["reward"], rew) and torch.equal(h["done"], dn), (t, side)
        assert torch.equal(ha.to_move, hc.to_move)


def test_ops_validate_their_tensors(ops):
    from hex_gym_env_b200 import HexBatch
    env = HexBatch(5, 100, variant=1, device=0, agent_mode=2)
    env.reset()
    good = _bufs(100, 5)
    with pytest.raises(RuntimeError, match="elements"):
        ops.step(env.handle, None, None, good["obs"][:50], None, None, None, None, None)
    with pytest.raises(RuntimeError, match="dtype"):
        ops.step(env.handle, None, None, good["obs"].to(torch.uint8), None, None, None, None, None)
    with pytest.raises(RuntimeError, match="cuda"):
        ops.step(env.handle, None, None, good["obs"].cpu(), None, None, None, None, None)
    with pytest.raises(RuntimeError, match="contiguous"):
        ops.step(env.handle, None, None, None, torch.empty((25, 100), dtype=torch.uint8, device="cuda").t(), None, None, None, None)
    with pytest.raises(RuntimeError, match="bad argument"):
        ops.ply(env.handle, torch.zeros(100, dtype=torch.int32, device="cuda"), None)       # not a raw handle


def test_ops_capture_into_a_cuda_graph(ops):
    from hex_gym_env_b200 import HexBatch
    N, G, K = 7, 4096, 10
    a = HexBatch(N, G, variant=1, device=0, seed=6, agent_mode=2)
    c = HexBatch(N, G, variant=1, device=0, seed=6, agent_mode=2)
    a.reset(); c.reset()
    b = _bufs(G, N)
    s = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        ops.step(c.handle, None, None, b["obs"], b["mask"], b["reward"], b["done"], None, None)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(K):
                ops.step(c.handle, None, None, b["obs"], b["mask"], b["reward"], b["done"], None, None)
    g.replay()
    g.replay()
    torch.cuda.synchronize()
    for _ in range(1 + 2 * K):   # one eager step, then two replays of K captured steps (capturing does not execute)
        o = a.step()
    for k in ("obs", "mask", "reward", "done"):
        assert torch.equal(o[k], b[k]), k


def test_masked_sample_op(ops):
    G, C = 333, 36
    logits = torch.randn(G, C, device="cuda")
    mask = (torch.rand(G, C, device="cuda") < 0.6).to(torch.uint8)
    mask[:, 0] = 1
    u = torch.rand(G, dtype=torch.float64, device="cuda")
    from hex_gym_env_b200.rollout import masked_sample
    a0, lp0, en0 = masked_sample(logits, mask, u, want_entropy=True)
    a1, lp1, en1 = torch.empty(G, dtype=torch.int32, device="cuda"), torch.empty(G, device="cuda"), torch.empty(G, device="cuda")
    ops.masked_sample(logits, mask, u, a1, lp1, en1)
    assert torch.equal(a0, a1) and torch.equal(lp0, lp1) and torch.equal(en0, en1)
