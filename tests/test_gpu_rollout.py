"""GPU (-m gpu): the rollout feed. The masked categorical sampler is a floating-point kernel, so it
is compared with a plain torch fp32 reference of the same op (tolerances stated below); the collector is checked for the
invariants SB3's MaskableRolloutBuffer relies on."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("C", [9, 36, 121, 361])
def test_masked_sample_vs_torch(C):
    import torch
    from hex_gym_env_b200.rollout import masked_sample
    G = 4099
    g = torch.Generator(device="cuda"); g.manual_seed(C)
    logits = 3.0 * torch.randn(G, C, device="cuda", generator=g)
    mask = (torch.rand(G, C, device="cuda", generator=g) < 0.4)
    mask[:, 0] |= ~mask.any(1)                      # at least one legal cell per row ...
    mask[7] = False                                  # ... except one row that has none
    u = torch.rand(G, dtype=torch.float64, device="cuda", generator=g)
    u[11], u[12] = 0.0, 1.0 - 2.0 ** -53             # edge draws
    a, logp, ent = masked_sample(logits, mask, u, want_entropy=True)
    # torch fp32 reference of the same op
    ml = logits.masked_fill(~mask, float("-inf"))
    ref_logp_all = torch.log_softmax(ml, dim=-1)
    p = torch.exp(ref_logp_all)
    cdf = torch.cumsum(p.double(), dim=-1)
    rows = torch.arange(G, device="cuda")
    ok = torch.ones(G, dtype=torch.bool, device="cuda"); ok[7] = False
    assert int(a[7]) == -1 and float(logp[7]) == 0.0
    al = a.long().clamp(min=0)
    assert bool(mask[rows, al][ok].all())                                   # only legal cells are ever drawn
    # log-probability of the drawn action: |kernel - torch| <= 2e-5 (fp32 exp/log, different summation order)
    assert float((logp - ref_logp_all[rows, al])[ok].abs().max()) <= 2e-5
    # inverse-CDF property, tolerance 1e-5 on the CDF (fp32 accumulation): cdf[a-1] - tol <= u < cdf[a] + tol
    hi = cdf[rows, al]
    lo = torch.where(al > 0, cdf[rows, (al - 1).clamp(min=0)], torch.zeros_like(hi))
    # cells between a-1 and a that are illegal contribute 0, so lo may be taken at the previous cell directly
    assert bool(((u >= lo - 1e-5) & (u <= hi + 1e-5))[ok].all())
    ref_ent = -(p * ref_logp_all.masked_fill(~mask, 0.0)).sum(-1)
    assert float((ent - ref_ent)[ok].abs().max()) <= 1e-4
    # statistical check: empirical frequencies of one row follow its probabilities (chi-square style bound)
    row = 3
    reps = 20000
    uu = torch.rand(reps, dtype=torch.float64, device="cuda", generator=g)
    aa, _ = masked_sample(logits[row:row + 1].expand(reps, C).contiguous(), mask[row:row + 1].expand(reps, C).contiguous(), uu)
    freq = torch.bincount(aa.long(), minlength=C).double() / reps
    assert float((freq - p[row].double()).abs().max()) < 0.02


def test_collector_invariants():
    import torch
    from hex_gym_env_b200 import HexBatch, VARIANT_B, AGENT_RANDOM
    from hex_gym_env_b200.rollout import RolloutCollector
    N, G, T = 6, 512, 32
    env = HexBatch(N, G, variant=VARIANT_B, device=0, seed=1, agent_mode=AGENT_RANDOM)
    lin = torch.nn.Linear(N * N, N * N + 1).cuda()

    def policy(obs):
        y = lin(obs.flatten(1))
        return y[:, :-1], y[:, -1]

    col = RolloutCollector(env, T, seed=2)
    for _ in range(2):
        buf = col.collect(policy)
        taken = buf.action_masks[:-1].gather(2, buf.actions.long().unsqueeze(-1)).squeeze(-1)
        assert bool((taken == 1).all())                                       # every sampled action was legal ...
        assert int(env.stats()[5]) == 0                                       # ... so no episode ended on an illegal move
        assert bool(((buf.obs[:-1] == 0).flatten(2) == (buf.action_masks[:-1] == 1)).all())   # mask == empty cells of the obs
        assert bool((buf.episode_starts[1:] == buf.dones.float()).all())
        assert bool((buf.rewards[buf.dones == 0] == 0).all()) and bool((buf.rewards[buf.dones == 1].abs() == 1).all())
        # GAE recomputed naively for one game
        g = 5
        adv, last = np.zeros(T), 0.0
        v, r, es = buf.values[:, g].cpu().numpy(), buf.rewards[:, g].cpu().numpy(), buf.episode_starts[:, g].cpu().numpy()
        for t in reversed(range(T)):
            nt = 1.0 - es[t + 1]
            delta = r[t] + 0.99 * v[t + 1] * nt - v[t]
            last = delta + 0.99 * 0.95 * nt * last
            adv[t] = last
        assert np.allclose(adv, buf.advantages[:, g].cpu().numpy(), atol=1e-5)
    assert int(env.stats()[6]) == 2 * T * G


def test_collector_cuda_graph_rollout():
    """collect(use_graph=True): the rollout replayed from one CUDA graph keeps every invariant of the eager rollout, advances
    the environment by exactly T steps per call, follows in-place parameter updates, and never plays an illegal move."""
    import torch
    from hex_gym_env_b200 import HexBatch, VARIANT_B, AGENT_RANDOM
    from hex_gym_env_b200.rollout import RolloutCollector
    N, G, T = 6, 1024, 16
    env = HexBatch(N, G, variant=VARIANT_B, device=0, seed=3, agent_mode=AGENT_RANDOM)
    lin = torch.nn.Linear(N * N, N * N + 1).cuda()

    def policy(obs):
        y = lin(obs.flatten(1))
        return y[:, :-1], y[:, -1]

    col = RolloutCollector(env, T, seed=4)
    last_obs = None
    for it in range(5):
        buf = col.collect(policy, use_graph=True)
        torch.cuda.synchronize()
        if last_obs is not None:
            assert torch.equal(buf.obs[0], last_obs)                           # continues where the previous rollout stopped
        last_obs = buf.obs[-1].clone()
        taken = buf.action_masks[:-1].gather(2, buf.actions.long().unsqueeze(-1)).squeeze(-1)
        assert bool((taken == 1).all()) and int(env.stats()[5]) == 0
        assert bool(((buf.obs[:-1] == 0).flatten(2) == (buf.action_masks[:-1] == 1)).all())
        assert bool((buf.episode_starts[1:] == buf.dones.float()).all())
        assert int(env.stats()[6]) == (it + 1) * T * G
        logits, values = policy(buf.obs[3].float())
        assert torch.allclose(values, buf.values[3], atol=1e-5)                # the graph reads the CURRENT parameters
        assert torch.allclose(buf.returns, buf.advantages + buf.values[:-1], atol=1e-6)
        with torch.no_grad():
            lin.weight.mul_(0.5).add_(0.01 * (it + 1))                         # an "optimizer step", in place
    # different rollouts really were sampled (the generator advances across replays)
    assert col._graph is not None
    a1 = buf.actions.clone()
    buf = col.collect(policy, use_graph=True)
    torch.cuda.synchronize()
    assert not torch.equal(a1, buf.actions)
