"""GPU (-m gpu): the host-buffer forms of the step (hexb_step_host, its begin/end split and the bit-packed transport) against
the device-resident step, and hexb_gae against the eager PyTorch loop it replaces."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    return torch


def _pair(torch, N, G, variant, seed, **kw):
    from hex_gym_env_b200 import HexBatch
    a = HexBatch(N, G, variant=variant, device=0, seed=seed, **kw)
    b = HexBatch(N, G, variant=variant, device=0, seed=seed, **kw)
    a.reset(); b.reset()
    return a, b


@pytest.mark.parametrize("N,G,variant", [(3, 1, 1), (5, 77, 1), (7, 4096, 0), (11, 5000, 1), (19, 333, 1), (6, 128, 1)])
def test_step_host_packed_matches_device(torch_mod, N, G, variant):
    torch = torch_mod
    kw = dict(agent_mode=2) if variant == 1 else {}
    dev, host = _pair(torch, N, G, variant, 11, **kw)
    io = host.pinned_io()
    gen = torch.Generator(device="cuda"); gen.manual_seed(3)
    for t in range(N * N // 2 + 6):
        if t % 3 == 2:
            acts_host = None
            o = dev.step()
        else:
            acts = dev.sample_actions(torch.rand(G, dtype=torch.float64, device="cuda", generator=gen))
            if t % 5 == 0:
                acts = acts.clone(); acts[::7] = 0          # some illegal moves
            acts_host = acts.cpu()
            o = dev.step(acts)
        io["obs"].fill_(99); io["mask"].fill_(99)
        host.step_host_packed(acts_host, io)
        for k in ("obs", "mask", "reward", "done"):
            assert np.array_equal(io[k].numpy(), o[k].cpu().numpy()), (k, t)
    assert torch.equal(dev.stats(), host.stats())
    assert host._lib.hexb_host_threads() >= 1


@pytest.mark.parametrize("frac", [-1.0, 0.0, 0.3, 0.5, 0.97, 1.0])
def test_step_host_transports_match_device(torch_mod, frac):
    """hexb_step_host with every DMA / packed split (adaptive, pinned fractions, plain DMA) returns the device step's bytes."""
    torch = torch_mod
    for N, G in ((11, 8192 + 77), (7, 4096), (5, 5000)):
        dev, host = _pair(torch, N, G, 1, 4, agent_mode=2)
        host.set_host_transport(frac)
        io = host.pinned_io()
        for t in range(25):
            o = dev.step()
            io["obs"].fill_(99); io["mask"].fill_(99)
            host.step_host(None, io)
            for k in ("obs", "mask", "reward", "done"):
                assert np.array_equal(io[k].numpy(), o[k].cpu().numpy()), (N, k, t, frac)
        f = host.host_transport()
        assert 0.0 <= f <= 1.0 and (frac < 0 or f == frac)


def test_step_host_begin_end_matches_step_host(torch_mod):
    torch = torch_mod
    a, b = _pair(torch, 7, 3000, 1, 5, agent_mode=2)
    ioa, iob = a.pinned_io(), b.pinned_io()
    for t in range(30):
        a.step_host(None, ioa)
        b.step_host_begin(None, iob)
        x = sum(range(1000))          # the host is free between begin and end
        with pytest.raises(Exception):
            b.step_host_begin(None, iob)          # one step may be pending
        b.step_host_end()
        for k in ("obs", "mask", "reward", "done"):
            assert np.array_equal(ioa[k].numpy(), iob[k].numpy()), (k, t)
    with pytest.raises(Exception):
        b.step_host_end()                         # nothing pending


def test_step_host_float32_observations(torch_mod):
    torch = torch_mod
    from hex_gym_env_b200 import HexBatch
    a = HexBatch(6, 500, variant=1, device=0, seed=2, agent_mode=2)
    b = HexBatch(6, 500, variant=1, device=0, seed=2, agent_mode=2, obs_dtype=torch.float32)
    a.reset(); b.reset()
    ioa, iob = a.pinned_io(), b.pinned_io()
    assert iob["obs"].dtype == torch.float32
    for t in range(25):
        a.step_host(None, ioa); b.step_host(None, iob)
        assert np.array_equal(ioa["obs"].numpy().astype(np.float32), iob["obs"].numpy())
        assert np.array_equal(ioa["mask"].numpy(), iob["mask"].numpy())
    with pytest.raises(Exception):
        b.step_host_packed(None, iob)             # the packed transport carries int8 observations only


@pytest.mark.parametrize("T,G", [(1, 1), (7, 33), (128, 4096), (64, 1000)])
def test_gae_matches_torch_loop(torch_mod, T, G):
    """hexb_gae against the T-iteration eager PyTorch recurrence (the float32 reference, SB3 RolloutBuffer semantics): the kernel
    performs the same float32 operations in the same order, so the results agree to <= 1e-6 relative (in fact bit for bit)."""
    torch = torch_mod
    from hex_gym_env_b200.rollout import gae
    g = torch.Generator(device="cuda"); g.manual_seed(T * 1000 + G)
    rewards = (torch.randint(-1, 2, (T, G), device="cuda", generator=g)).float() * (torch.rand((T, G), device="cuda", generator=g) < 0.1)
    values = torch.randn((T + 1, G), device="cuda", generator=g)
    dones = (torch.rand((T, G), device="cuda", generator=g) < 0.07).to(torch.uint8)
    for gamma, lam in ((0.99, 0.95), (1.0, 1.0), (0.9, 0.0)):
        adv, ret = gae(rewards, values, dones, gamma, lam)
        last = torch.zeros(G, device="cuda")
        want = torch.empty_like(adv)
        for t in reversed(range(T)):
            nonterminal = 1.0 - dones[t].float()
            delta = rewards[t] + gamma * values[t + 1] * nonterminal - values[t]
            last = delta + gamma * lam * nonterminal * last
            want[t] = last
        assert torch.allclose(adv, want, rtol=1e-6, atol=1e-7), (T, G, gamma, lam, (adv - want).abs().max().item())
        assert torch.allclose(ret, want + values[:-1], rtol=1e-6, atol=1e-7)


def test_rollout_buffer_uses_the_kernel_and_agrees_with_the_loop(torch_mod):
    torch = torch_mod
    from hex_gym_env_b200 import HexBatch
    from hex_gym_env_b200.rollout import RolloutCollector
    b = HexBatch(6, 512, variant=1, device=0, seed=1, agent_mode=2, obs_dtype=torch.float32)
    net = torch.nn.Linear(36, 37).cuda()

    def policy(obs):
        y = net(obs.reshape(obs.shape[0], -1))
        return y[:, :36], y[:, 36]
    col = RolloutCollector(b, 32, seed=3)
    buf = col.collect(policy)
    assert buf.obs.dtype == torch.float32
    adv, ret = buf.advantages.clone(), buf.returns.clone()
    buf.compute_returns_and_advantage_torch()
    assert torch.allclose(adv, buf.advantages, rtol=1e-6, atol=1e-7) and torch.allclose(ret, buf.returns, rtol=1e-6, atol=1e-7)


def test_checkpoint_roundtrip_manual_opponent(torch_mod):
    """state_dict carries the per-game opponent bookkeeping that lives outside the packed blob (opp_index, to_move, info)."""
    torch = torch_mod
    from hex_gym_env_b200 import HexBatch
    kw = dict(variant=1, device=0, seed=9, agent_mode=2, manual_opponent=True, pool_size=5)
    a = HexBatch(5, 300, **kw)
    a.reset()
    a.opponent_catch_up()
    gen = torch.Generator(device="cuda"); gen.manual_seed(1)
    for t in range(12):
        acts = a.sample_actions(torch.rand(300, dtype=torch.float64, device="cuda", generator=gen))
        a.half_step(0, acts); a.half_step(1, None); a.half_step(1, None)
    sd = a.state_dict()
    assert "opp_index" in sd and "to_move" in sd
    assert (sd["opp_index"] >= 0).any()        # some games play a pool opponent: not the constructor's -1
    b = HexBatch(5, 300, **kw)
    b.load_state_dict(sd)
    assert torch.equal(b.opp_index, a.opp_index) and torch.equal(b.to_move, a.to_move)
    for t in range(10):
        acts = a.sample_actions(torch.rand(300, dtype=torch.float64, device="cuda", generator=gen))
        for env in (a, b):
            env.half_step(0, acts); env.half_step(1, None); env.half_step(1, None)
        assert torch.equal(a.opp_index, b.opp_index) and torch.equal(a.to_move, b.to_move)
    ea, eb = a.export_state(), b.export_state()
    for k in ea:
        assert torch.equal(ea[k], eb[k]), k
