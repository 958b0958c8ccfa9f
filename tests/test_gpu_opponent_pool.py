"""GPU (-m gpu): the opponent pool end to end - HexBatch.step_with_opponent driven by an OpponentPool of scripted models against
the unmodified reference run with OpponentPolicy opponents (tests/golden/oppmodel_*.npz, evalpool_*.npz: the latter switch
SelfPlayEnv.set_eval on and off in mid-run), the evaluation pass over the pool, and the checkpoint of the evaluation counters."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_files
from oracle.scripted import scripted_choice

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")


class Scripted(object):
    """The rule of oracle/scripted.py::ScriptedModel as a batched policy (evaluated on the host: this is a test double)."""

    def __init__(self):
        self.games = 0

    def __call__(self, obs, mask):
        o, m = obs.cpu().numpy(), mask.cpu().numpy()
        self.games += len(o)
        a = [scripted_choice(o[g], m[g]) if m[g].any() else 0 for g in range(len(o))]
        return torch.tensor(a, dtype=torch.int32, device=obs.device)


def _pool_and_batch(N, G, seed, agent_mode, pool_size, dense):
    from hex_gym_env_b200 import VARIANT_B, HexBatch
    from hex_gym_env_b200.opponents import OpponentPool
    b = HexBatch(N, G, variant=VARIANT_B, device=0, seed=seed, agent_mode=agent_mode, manual_opponent=True, pool_size=pool_size)
    pool = OpponentPool(Scripted(), buffer_size=pool_size, batch=b, dense=dense)
    for k in range(pool_size):
        pool.set_opponent_model(k, Scripted(), 0.0)          # distinct objects: every entry is dispatched on its own
    return b, pool


@pytest.mark.parametrize("dense", [True, False])
@pytest.mark.parametrize("name", golden_files("evalpool_") + ["oppmodel_N7_a2.npz"])
def test_step_with_opponent_pool_against_the_reference(name, dense):
    z = np.load(os.path.join(GOLDEN, name))
    N, seed, am, K = int(z["N"]), int(z["seed"]), int(z["agent_mode"]), int(z["pool"])
    T, G = z["actions"].shape
    eval_at = z["eval_at"] if "eval_at" in z.files else None
    b, pool = _pool_and_batch(N, G, seed, am, K, dense)
    b.reset()
    b.opponent_opening(pool)
    obs, mask = b.encode(0)
    assert np.array_equal(obs.cpu().numpy(), z["obs0"]) and np.array_equal(mask.cpu().numpy(), z["mask0"])
    for t in range(T):
        if eval_at is not None and eval_at[t] >= 0:
            pool.set_eval(bool(eval_at[t]))
            assert b.eval_state == bool(eval_at[t])
        b._buf("sw_term", (G, N, N), b.obs_dtype).zero_()
        out = b.step_with_opponent(torch.from_numpy(z["actions"][t]).cuda(), pool, want_term=True)
        done = z["done"][t].astype(bool)
        w = "%s t=%d" % (name, t)
        assert np.array_equal(out["done"].cpu().numpy().astype(bool), done), w
        assert np.array_equal(out["reward"].cpu().numpy(), z["reward"][t]), w
        assert np.array_equal(out["obs"].cpu().numpy(), z["obs"][t]), w
        assert np.array_equal(out["mask"].cpu().numpy(), z["mask"][t]), w
        assert np.array_equal(out["term_obs"].cpu().numpy()[done], z["term_obs"][t][done]), w
    if not dense:   # the gather form never shows a model a game that is not its own
        shown = sum(m.games for m, _ in pool.groups())
        asked = int((z["opp_model"] != -9).sum()) + int((z["opp0_model"] != -9).sum())
        assert shown == asked, (shown, asked)


def test_evaluate_pool_meets_every_entry_once_per_game():
    from hex_gym_env_b200.opponents import evaluate_pool
    from hex_gym_env_b200.rollout import masked_sample
    N, G, K = 5, 300, 4
    b, pool = _pool_and_batch(N, G, 11, 2, K, dense=True)
    gen = torch.Generator(device="cuda").manual_seed(5)

    def random_agent(obs, mask):
        return masked_sample(torch.zeros(mask.shape, dtype=torch.float32, device=mask.device), mask, generator=gen)[0]

    before = b.stats().cpu().numpy().copy()
    r = evaluate_pool(b, pool, random_agent)
    assert r["episodes"] == G * K and list(r["per_entry_episodes"]) == [G] * K
    assert -1.0 <= r["mean_reward"] <= 1.0 and r["per_entry"].shape == (K,)
    assert abs(r["mean_reward"] - float(np.mean(r["per_entry"]))) < 1e-12      # equal counts: mean of means
    assert not b.eval_state and not pool.eval_state
    assert (b.stats().cpu().numpy() - before)[0] >= G * K                      # the device counted at least those episodes
    assert int((b.to_move == 0).all())                                         # left reset, the agent to move everywhere


def test_checkpoint_carries_the_evaluation_cycle():
    N, G, K = 4, 64, 5
    a, pool_a = _pool_and_batch(N, G, 3, 2, K, dense=True)
    gen = torch.Generator(device="cuda").manual_seed(1)

    def play(batch, pool, steps, g):
        outs = []
        obs, mask = batch.encode(0)
        for _ in range(steps):
            u = torch.rand(G, generator=g, device="cuda", dtype=torch.float64)
            acts = batch.sample_actions(u, 0)
            o = batch.step_with_opponent(acts, pool)
            outs.append((o["obs"].clone(), o["reward"].clone(), o["done"].clone(), batch.opp_index.clone()))
        return outs

    a.reset()
    a.opponent_opening(pool_a)
    pool_a.set_eval(True)
    play(a, pool_a, 12, gen)
    sd = a.state_dict()
    assert "eval_episode" in sd and sd["config"]["eval_state"] == 1 and int(sd["eval_episode"].max()) >= 1
    gstate = gen.get_state()
    want = play(a, pool_a, 25, gen)
    b, pool_b = _pool_and_batch(N, G, 3, 2, K, dense=True)     # a fresh object in training mode
    b.load_state_dict(sd)
    assert b.eval_state
    pool_b.eval_state = True
    gen.set_state(gstate)
    got = play(b, pool_b, 25, gen)
    for t, (w, g) in enumerate(zip(want, got)):
        for x, y in zip(w, g):
            assert torch.equal(x, y), t


@pytest.mark.parametrize("name", ["evalpool_N5_a2.npz", "oppmodel_N4_a1.npz"])
def test_vec_env_with_an_opponent_pool_against_the_reference(name):
    """HexVecEnv(base_model=..., buffer_size=...) = selfplay_wrapper(HexEnv)(base_model=..., buffer_size=...) for every game: the
    SB3-shaped surface on top of the pool, set_eval through env_method like SelfPlayCallback reaches it."""
    from hex_gym_env_b200.vec_env import HexVecEnv
    z = np.load(os.path.join(GOLDEN, name))
    N, seed, am, K = int(z["N"]), int(z["seed"]), int(z["agent_mode"]), int(z["pool"])
    T, G = z["actions"].shape
    eval_at = z["eval_at"] if "eval_at" in z.files else None
    env = HexVecEnv(board_size=N, num_envs=G, seed=seed, agent_player_num=None if am == 2 else am, base_model=Scripted(),
                    buffer_size=K, scores=np.zeros(K), device=0)
    assert len(env.get_opponent_models()) == K and list(env.get_scores()) == [0.0] * K and env.best_score == 0.0
    for k in range(K):
        env.set_opponent_model(k, Scripted(), 0.0)
    obs = env.reset()
    assert obs.dtype == np.float32 and np.array_equal(obs, z["obs0"].astype(np.float32))
    assert np.array_equal(np.stack(env.env_method("action_masks")), z["mask0"].astype(bool))
    for t in range(T):
        if eval_at is not None and eval_at[t] >= 0:
            r = env.env_method("set_eval", bool(eval_at[t]))
            assert len(r) == G and env.eval_state == bool(eval_at[t]) and env.batch.eval_state == bool(eval_at[t])
        obs, rew, done, infos = env.step(z["actions"][t])
        w = "%s t=%d" % (name, t)
        assert np.array_equal(done, z["done"][t].astype(bool)), w
        assert np.array_equal(rew, z["reward"][t]), w
        assert np.array_equal(obs, z["obs"][t].astype(np.float32)), w
        assert np.array_equal(env.action_masks(), z["mask"][t].astype(bool)), w
        for g in np.flatnonzero(done):
            assert np.array_equal(infos[g]["terminal_observation"], z["term_obs"][t][g].astype(np.float32)), w
    with pytest.raises(AttributeError):
        HexVecEnv(board_size=4, num_envs=8, device=0).get_scores()      # no pool: the pool API is absent, not silently empty
    env.close()


def test_stacked_pool_samples_and_lives_in_the_rollout_graph():
    """StackedMlpOpponents on the device: its sampled actions are hexb_masked_sample of its gathered logits with the same
    uniforms, and a pool made of its slots runs inside RolloutCollector's CUDA graph, which is captured again when the pool changes."""
    import torch.nn as nn
    from hex_gym_env_b200 import VARIANT_B, HexBatch, OpponentPool, StackedMlpOpponents
    from hex_gym_env_b200.rollout import RolloutCollector, masked_sample
    N, G, K, T = 4, 512, 3, 8
    C = N * N
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)

    class Policy(nn.Module):
        def __init__(self):
            super().__init__()
            self.pi = nn.Sequential(nn.Flatten(), nn.Linear(C, 32), nn.Tanh(), nn.Linear(32, 32), nn.Tanh(), nn.Linear(32, C))
            self.vf = nn.Sequential(nn.Flatten(), nn.Linear(C, 32), nn.Tanh(), nn.Linear(32, 1))

        def forward(self, obs):
            return self.pi(obs), self.vf(obs).squeeze(-1)

    policy = Policy().to(dev)
    linears = lambda: [m for m in policy.pi if isinstance(m, nn.Linear)]
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    stack = StackedMlpOpponents((C, 32, 32, C), K + 1, device=dev, generator=gen)
    for sl in range(K + 1):
        stack.load(sl, linears())
    obs = torch.randint(-1, 2, (G, N, N), device=dev).to(torch.int8)
    mask = (obs.reshape(G, C) == 0).to(torch.uint8)
    mask[:, 3] = 1
    slots = torch.randint(0, K + 1, (G,), device=dev)
    with torch.no_grad():
        assert torch.allclose(stack.logits(obs, slots), policy.pi(obs.float()), atol=1e-5)     # every slot holds the same network here
    state = gen.get_state()
    got = stack.actions(obs, mask, slots)
    gen.set_state(state)
    u = torch.rand(G, dtype=torch.float64, device=dev, generator=gen)
    assert torch.equal(got, masked_sample(stack.logits(obs, slots), mask, u=u)[0])
    assert bool((mask.gather(1, got.long().unsqueeze(1)) == 1).all())

    env = HexBatch(N, G, variant=VARIANT_B, device=0, seed=2, agent_mode=2, manual_opponent=True, pool_size=K, obs_dtype=torch.float32)
    pool = OpponentPool(stack.entry(0), buffer_size=K, batch=env)
    for k in range(K):
        pool.set_opponent_model(k, stack.entry(k + 1), 0.0)
    col = RolloutCollector(env, T, seed=3, extra_generators=[gen])
    col.collect(policy, pool, use_graph=True)        # the first rollout runs eagerly
    col.collect(policy, pool, use_graph=True)        # captured and replayed
    key = col._graph_key
    col.collect(policy, pool, use_graph=True)        # replayed
    assert col._graph_key == key
    with torch.no_grad():
        policy.pi[1].weight.mul_(1.5)
    stack.load(2, linears())                         # the learner's weights into slot 2 (entry 1), which becomes the best model
    pool.set_opponent_model(1, stack.entry(2), 0.5)
    col.collect(policy, pool, use_graph=True)
    assert col._graph_key != key and pool.best_model.slot == 2
    torch.cuda.synchronize()
    st = env.stats().cpu().numpy()
    assert st[0] > 0 and st[5] == 0                  # episodes finished, none by an illegal agent move
    buf = col.buf
    assert bool(torch.isfinite(buf.advantages).all()) and bool((buf.action_masks[:-1].gather(2, buf.actions.long().unsqueeze(2)) == 1).all())


class ScriptedA(object):
    """oracle/scripted.py::ScriptedModelA's rule as a batched policy for the variant-A opponent view (a host-side test double)."""

    def __call__(self, obs, mask):
        from oracle.scripted import scripted_choice_a
        o, m = obs.cpu().numpy(), mask.cpu().numpy()
        return torch.tensor([scripted_choice_a(o[g], m[g]) if m[g].any() else 0 for g in range(len(o))], dtype=torch.int32, device=obs.device)


@pytest.mark.parametrize("name", golden_files("oppredict_"))
def test_vec_env_hex_v0_with_opponent_predict_against_the_reference(name):
    """HexVecEnv(variant="hex-v0", opponent_model=..., eps=...) = gym.make("hex-v0", opponent_policy="opponent_predict", ...) for
    every game (scripts/selfplay.py:38-44), against the unmodified reference run one env per game."""
    from hex_gym_env_b200.vec_env import HexVecEnv
    z = np.load(os.path.join(GOLDEN, name))
    N, seed, eps, of = int(z["N"]), int(z["seed"]), float(z["eps"]), int(z["opponent_first"])
    T, G = z["actions"].shape
    env = HexVecEnv(board_size=N, num_envs=G, variant="hex-v0", seed=seed, opponent_first=bool(of), opponent_model=ScriptedA(),
                    eps=eps, device=0)
    obs = env.reset()
    assert np.array_equal(obs, z["obs0"].astype(np.float32))
    for t in range(T):
        obs, rew, done, infos = env.step(z["actions"][t])
        w = "%s t=%d" % (name, t)
        assert np.array_equal(done, z["done"][t].astype(bool)), w
        assert np.array_equal(rew, z["reward"][t]), w
        assert np.array_equal(obs, z["obs"][t].astype(np.float32)), w
        assert np.array_equal(env.action_masks(), z["obs"][t].reshape(G, -1) == 2), w
    assert np.array_equal(env.batch.export_state()["draws"].cpu().numpy().astype(np.uint32), z["draws"][T - 1])
    other = ScriptedA()
    env.env_method("set_opponent_model", other)
    assert env.opponent_model is other
    with pytest.raises(ValueError):
        HexVecEnv(board_size=4, num_envs=8, opponent_model=ScriptedA(), device=0)     # SelfPlayEnv takes base_model instead
    env.close()


def test_vec_env_info_dict_of_hex_v0():
    """HexVecEnv(variant="hex-v0", info_fields=True): infos carry last_move_opponent / last_move_player / winner of the reference's
    info dict (HexGame.py:281-286), checked against the oracle's env-level bookkeeping on the same seeded run."""
    from hex_gym_env_b200.vec_env import HexVecEnv
    from oracle import hexref
    N, G, T, seed = 5, 200, 40, 21
    C = N * N
    env = HexVecEnv(board_size=N, num_envs=G, variant="hex-v0", seed=seed, info_fields=True, device=0)
    ref = hexref.RefBatch(hexref.KIND_ENV_A, N, G, seed=seed)
    obs = env.reset()
    robs, rmask = ref.reset()
    assert np.array_equal(obs, robs.astype(np.float32))
    rs = np.random.RandomState(4)
    seen_win = seen_illegal = 0
    for t in range(T):
        mask = env.action_masks()
        assert np.array_equal(mask, rmask.astype(bool))
        k = (rs.rand(G) * np.maximum(mask.sum(1), 1)).astype(np.int64)
        acts = np.argsort(-mask.astype(np.int8), axis=1, kind="stable")[np.arange(G), k].astype(np.int32)
        bad = rs.rand(G) < 0.03
        acts[bad] = rs.randint(0, C, size=int(bad.sum()))
        obs, rew, done, infos = env.step(acts)
        r = ref.step(acts, want_term=True)
        fo, fw = ref.info()
        assert np.array_equal(rew, r["reward"]) and np.array_equal(done, r["done"].astype(bool))
        for i in range(G):
            assert infos[i]["last_move_player"] == int(acts[i])
            assert infos[i]["last_move_opponent"] == (None if fo[i] < 0 else int(fo[i])), (t, i)
            assert infos[i]["winner"] == (None if fw[i] < 0 else int(fw[i])), (t, i)
            assert ("terminal_observation" in infos[i]) == bool(done[i])
        seen_win += int((fw >= 0).sum() - (fw == 3).sum())
        seen_illegal += int((fw == 3).sum())
        rmask = r["mask"]
    assert seen_win > 0 and seen_illegal > 0
    with pytest.raises(ValueError):
        HexVecEnv(board_size=4, num_envs=8, info_fields=True, device=0)       # SelfPlayEnv's step returns an empty info dict
    env.close()
