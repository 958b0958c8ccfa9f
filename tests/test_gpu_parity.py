"""GPU (-m gpu): the CUDA library (libhexb.so) called through its C ABI, against the golden vectors of the unmodified
reference and against the oracle on the same seeded inputs. Bit-exact on every array (integer / byte work)."""
import numpy as np
import pytest

import parity
from conftest import golden_files
from oracle import hexref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def make():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    from gpu_adapter import make as mk
    return mk


@pytest.mark.parametrize("name", golden_files("game_A"))
def test_golden_raw_games(make, name):
    parity.golden_raw_game(make, name)


@pytest.mark.parametrize("name", golden_files("selfplay_") + golden_files("envA_"))
def test_golden_rollouts(make, name):
    parity.golden_rollout(make, name)


@pytest.mark.parametrize("N", [3, 4, 5, 6, 7, 8, 9, 11, 13, 16, 19])
@pytest.mark.parametrize("agent_mode", [0, 1, 2])
def test_selfplay_vs_oracle(make, N, agent_mode):
    G = 1000 if N <= 11 else 300   # not a multiple of the 128-game tile: exercises the ragged last tile
    T = 2 * N * N // 3 + 10
    parity.versus_oracle(make, hexref.KIND_SELFPLAY_B, N, G, T, seed=N * 10 + agent_mode, game_offset=12345678901 * agent_mode,
                         fused=(agent_mode != 1), agent_mode=agent_mode, check_state_every=3)


@pytest.mark.parametrize("N", [3, 5, 7, 10, 11])
@pytest.mark.parametrize("opponent_first", [False, True])
def test_envA_vs_oracle(make, N, opponent_first):
    parity.versus_oracle(make, hexref.KIND_ENV_A, N, 700, N * N, seed=5 + N, fused=not opponent_first, opponent_first=opponent_first,
                         check_state_every=3)


@pytest.mark.parametrize("kind", [hexref.KIND_SELFPLAY_B, hexref.KIND_ENV_A])
def test_no_auto_reset(make, kind):
    kw = dict(agent_mode=2) if kind == hexref.KIND_SELFPLAY_B else {}
    parity.versus_oracle(make, kind, 4, 500, 30, seed=3, fused=False, auto_reset=False, illegal_rate=0.1, **kw)


def test_eval_state_draws(make):
    parity.versus_oracle(make, hexref.KIND_SELFPLAY_B, 5, 130, 40, seed=9, agent_mode=2, eval_state=True)


def test_config2_subset(make):
    """BASELINE config 2 (7x7 variant-A HexEnv with action masking): a 4,096-game subset against the oracle."""
    parity.versus_oracle(make, hexref.KIND_ENV_A, 7, 4096, 60, seed=2, fused=True, check_state_every=10)


def test_config3_subset(make):
    """BASELINE config 3 (11x11 SelfPlayEnv, random opponent, agent colour random per game): 8,192 games, 130 steps."""
    parity.versus_oracle(make, hexref.KIND_SELFPLAY_B, 11, 8192, 130, seed=0, fused=True, agent_mode=2, check_state_every=13)


def test_sharding_invariance_same_device(make):
    """Games keyed by GLOBAL index: one batch of 1000 games == two shards (offsets 0 and 437) of the same global batch."""
    N, G, T, cut = 7, 1000, 50, 437
    whole = make(hexref.KIND_SELFPLAY_B, N, G, seed=21, agent_mode=2)
    a = make(hexref.KIND_SELFPLAY_B, N, cut, seed=21, agent_mode=2)
    b = make(hexref.KIND_SELFPLAY_B, N, G - cut, seed=21, game_offset=cut, agent_mode=2)
    whole.reset(); a.reset(); b.reset()
    for t in range(T):
        w, x, y = whole.step(), a.step(), b.step()
        for k in ("obs", "mask", "reward", "done", "actions"):
            assert np.array_equal(w[k], np.concatenate([x[k], y[k]])), (k, t)
    assert np.array_equal(whole.stats(), a.stats() + b.stats())


def test_vec_env_surface():
    """HexVecEnv (SB3 VecEnv shape): numpy and torch outputs agree with the oracle, auto-reset + terminal_observation."""
    import torch
    from hex_gym_env_b200.vec_env import HexVecEnv
    N, G, T = 6, 257, 60
    ref = hexref.RefBatch(hexref.KIND_SELFPLAY_B, N, G, seed=5, agent_mode=2)
    venv = HexVecEnv(board_size=N, num_envs=G, variant="selfplay", seed=5, output="numpy")
    tenv = HexVecEnv(board_size=N, num_envs=G, variant="selfplay", seed=5, output="torch")
    robs, rmask = ref.reset()
    obs = venv.reset()
    tobs = tenv.reset()
    assert obs.dtype == np.float32 and np.array_equal(obs, robs.astype(np.float32)) and np.array_equal(tobs.cpu().numpy(), obs)
    rs = np.random.RandomState(0)
    for t in range(T):
        masks = venv.action_masks()
        assert masks.dtype == bool and np.array_equal(masks, rmask.astype(bool))
        assert np.array_equal(np.stack(venv.env_method("action_masks")), masks)
        cnt = masks.sum(1)
        k = (rs.rand(G) * cnt).astype(np.int64)
        acts = np.argsort(~masks, axis=1, kind="stable")[np.arange(G), k]
        r = ref.step(acts.astype(np.int32), want_term=True)
        obs, rew, dones, infos = venv.step(acts)
        tobs, trew, tdones, tinfos = tenv.step(torch.as_tensor(acts, dtype=torch.int32, device="cuda"))
        assert np.array_equal(obs, r["obs"].astype(np.float32)) and np.array_equal(rew, r["reward"])
        assert np.array_equal(dones, r["done"].astype(bool)) and len(infos) == G
        assert np.array_equal(tobs.cpu().numpy(), obs) and np.array_equal(tdones.cpu().numpy(), dones)
        for i in np.flatnonzero(dones):
            assert np.array_equal(infos[i]["terminal_observation"], r["term_obs"][i].astype(np.float32))
            assert np.array_equal(tinfos[i]["terminal_observation"].cpu().numpy(), r["term_obs"][i])
        for i in np.flatnonzero(~dones)[:3]:
            assert infos[i] == {}
        rmask = r["mask"]
    st = venv.episode_stats()
    assert st["episodes"] == int(ref.stats()[0]) and st["env_steps"] == G * T


def test_step_host_matches_device(make):
    """hexb_step_host (HOST buffers in and out) == hexb_step."""
    from hex_gym_env_b200 import HexBatch, VARIANT_B
    N, G = 5, 300
    a = HexBatch(N, G, variant=VARIANT_B, device=0, seed=3, agent_mode=2)
    b = HexBatch(N, G, variant=VARIANT_B, device=0, seed=3, agent_mode=2)
    a.reset(); b.reset()
    for t in range(30):
        o = a.step(want_actions=True)
        io = b.step_host(None)
        for k in ("obs", "mask", "reward", "done"):
            assert np.array_equal(o[k].cpu().numpy(), io[k].numpy()), (k, t)
    acts = a.sample_actions(np.full(G, 0.5))
    o = a.step(acts)
    io = b.step_host(acts.cpu())
    for k in ("obs", "mask", "reward", "done"):
        assert np.array_equal(o[k].cpu().numpy(), io[k].numpy()), k


@pytest.mark.parametrize("N", [7, 11, 19])
def test_snake_chain_worst_case(make, N):
    parity.snake_chain(make, N)


@pytest.mark.parametrize("name", golden_files("oppmodel_") + golden_files("evalpool_"))
def test_golden_scripted_opponent(make, name):
    """hexb_half_step (learned opponent) against the reference run with OpponentPolicy opponents."""
    parity.golden_oppmodel(make, name)


@pytest.mark.parametrize("N,variant_kind", [(5, hexref.KIND_SELFPLAY_B), (11, hexref.KIND_SELFPLAY_B), (6, hexref.KIND_ENV_A)])
def test_half_step_vs_oracle(make, N, variant_kind):
    """Random legal (and a few illegal) moves of both sides through the split step, GPU vs oracle, with and without auto-reset."""
    G, T, C = 700, 2 * N * N // 3, N * N
    for auto_reset in (True, False):
        kw = dict(agent_mode=2) if variant_kind == hexref.KIND_SELFPLAY_B else dict(opponent_first=True)
        env = make(variant_kind, N, G, seed=N, auto_reset=auto_reset, manual_opponent=True, pool_size=7, **kw)
        ref = hexref.RefBatch(variant_kind, N, G, seed=N, manual_opponent=True, pool_size=7, **kw)
        env.reset(); ref.reset()
        if variant_kind == hexref.KIND_ENV_A:     # HexEnv.opponent_predict's eps mix: the draw decides who moves, on both sides alike
            env.set_opponent_eps(0.35); ref.set_opponent_eps(0.35)
        rs = np.random.RandomState(N)
        for t in range(T):
            if variant_kind == hexref.KIND_SELFPLAY_B and t in (T // 4, (2 * T) // 3):   # SelfPlayEnv.set_eval on, later off again
                env.set_eval(t == T // 4); ref.set_eval(t == T // 4)
            for side in (1, 0, 1):
                tm, idx = env.opp_state()
                rtm, ridx = ref.opp_state()
                parity.eq(tm, rtm, "to_move t=%d" % t)
                parity.eq(idx, ridx, "opp_index t=%d" % t)
                obs1, mask1 = env.view1()
                # (variant A: the opponent sees the board transposed with the colours swapped, HexGame.py:333-339 - the oracle's
                # view1 applies the reference's own invert_board; checked here independently of it as well)
                robs1, rmask1 = ref.view1()
                live = tm != 2
                parity.eq(obs1[live], robs1[live], "side-to-move obs t=%d" % t)
                parity.eq(mask1[live], rmask1[live], "side-to-move mask t=%d" % t)
                if variant_kind == hexref.KIND_ENV_A:
                    true_board = ref.observe()[0]
                    oppv = tm == 1
                    want = np.where(true_board[oppv] == 2, 2, 1 - true_board[oppv]).transpose(0, 2, 1)
                    parity.eq(obs1[oppv], want, "opponent view (A) t=%d" % t)
                cnt = np.maximum(mask1.sum(1), 1)
                k = (rs.rand(G) * cnt).astype(np.int64)
                acts = np.argsort(-mask1.astype(np.int8), axis=1, kind="stable")[np.arange(G), k].astype(np.int32)
                bad = rs.rand(G) < 0.02
                acts[bad] = rs.randint(-1, C + 1, size=int(bad.sum()))
                o = env.half_step(side, acts, want_term=True)
                r = ref.half_step(side, acts, auto_reset=auto_reset, want_term=True)
                for key in ("reward", "done", "to_move", "opp_index"):
                    parity.eq(o[key], r[key], "%s t=%d side=%d" % (key, t, side))
                d = r["done"].astype(bool) if auto_reset else np.zeros(G, bool)
                parity.eq(o["term_obs"][d], r["term_obs"][d], "term_obs t=%d" % t)
            if t % 5 == 0:
                e, re_ = env.export(), ref.export()
                for key in ("regions", "region_counter", "cur", "done", "winner", "agent", "draws"):
                    parity.eq(e[key], re_[key], "%s t=%d" % (key, t))
        parity.eq(env.stats(), ref.stats(), "stats")


def test_step_with_opponent_on_device():
    """HexBatch.step_with_opponent with the scripted rule as a torch callable == three explicit half steps."""
    import torch
    from hex_gym_env_b200 import HexBatch, VARIANT_B
    N, G = 7, 513
    a = HexBatch(N, G, variant=VARIANT_B, device=0, seed=4, agent_mode=2, manual_opponent=True, pool_size=4)
    b = HexBatch(N, G, variant=VARIANT_B, device=0, seed=4, agent_mode=2, manual_opponent=True, pool_size=4)

    def first_legal(obs, mask, to_move, opp_index):
        return torch.argmax(mask, dim=1).to(torch.int32)

    for env in (a, b):
        env.reset()
        o1, m1 = env.encode(1)
        env.half_step(1, first_legal(o1, m1, None, None))
    gen = torch.Generator(device="cuda"); gen.manual_seed(0)
    for t in range(60):
        _, mask = a.encode(0)
        acts = a.sample_actions(torch.rand(G, dtype=torch.float64, device="cuda", generator=gen)).clone()
        out = a.step_with_opponent(acts, first_legal)
        h = b.half_step(0, acts)
        rew, done = h["reward"].clone(), h["done"].clone()
        for _ in range(2):
            o1, m1 = b.encode(1)
            h = b.half_step(1, first_legal(o1, m1, None, None))
            rew += h["reward"]; done |= h["done"]
        obs, mask = b.encode(0)
        assert torch.equal(out["obs"], obs) and torch.equal(out["mask"], mask) and torch.equal(out["reward"], rew) and torch.equal(out["done"], done)
        assert bool((a.to_move == 0).all())
    assert int(a.stats()[0]) > 0 and int(a.stats()[5]) == 0


@pytest.mark.parametrize("N,kind,kw", [(5, hexref.KIND_SELFPLAY_B, dict(agent_mode=2)), (11, hexref.KIND_SELFPLAY_B, dict(agent_mode=2)),
                                       (7, hexref.KIND_ENV_A, dict(opponent_first=True)), (19, hexref.KIND_SELFPLAY_B, dict(agent_mode=0))])
def test_rollout_equals_steps(make, N, kind, kw):
    parity.rollout_equals_steps(make, kind, N, 1000 if N < 19 else 300, min(N * N // 2 + 3, 64), seed=N, **kw)


def test_checkpoint_roundtrip():
    """state_dict / load_state_dict: a restored shard continues bit-identically (state, random streams, statistics)."""
    import torch
    from hex_gym_env_b200 import HexBatch, VARIANT_B
    a = HexBatch(7, 777, variant=VARIANT_B, device=0, seed=8, agent_mode=2)
    a.reset()
    for _ in range(25):
        a.step()
    sd = a.state_dict()
    ref = [{k: v.clone() for k, v in a.step().items()} for _ in range(20)]
    b = HexBatch(7, 777, variant=VARIANT_B, device=0, seed=8, agent_mode=2)
    b.load_state_dict(sd)
    for t in range(20):
        o = b.step()
        for k in ("obs", "mask", "reward", "done"):
            assert torch.equal(o[k], ref[t][k]), (k, t)
    assert torch.equal(a.stats(), b.stats())
    with pytest.raises(ValueError):
        HexBatch(7, 778, variant=VARIANT_B, device=0, seed=8, agent_mode=2).load_state_dict(sd)


@pytest.mark.parametrize("N,agent_mode", [(5, 2), (11, 1), (11, 2)])
def test_sample_board_flow(make, N, agent_mode):
    """Batched sample_board=True (random start positions + random opponent) through import + split steps, GPU vs oracle."""
    parity.sample_board_flow(make, N, 600, 50, seed=N + agent_mode, agent_mode=agent_mode)


@pytest.mark.parametrize("name", golden_files("preset_"))
def test_golden_preset_boards(name):
    """hexb_import_boards on raw handles against the reference's HexGame.__init__ with a preset board."""
    import os
    from conftest import GOLDEN
    from gpu_adapter import GpuBatch

    def make_raw(kind, N, G):
        z = np.load(os.path.join(GOLDEN, name))
        env = GpuBatch(0 if kind == hexref.KIND_GAME_A else 1, N, G, raw=True)
        env.reset()
        env.import_boards(z["board_true"], np.zeros(G, np.int8))
        return env
    parity.golden_preset(make_raw, name)


def test_vec_env_sample_board():
    """HexVecEnv(sample_board=True): episodes start from random even positions; invariants of the reference's random_board."""
    import torch
    from hex_gym_env_b200.vec_env import HexVecEnv, random_start_boards
    N, G = 8, 500
    gen = torch.Generator(device="cuda"); gen.manual_seed(1)
    bd = random_start_boards(4000, N, gen, torch.device("cuda", 0))
    nb, nw = (bd == 0).flatten(1).sum(1), (bd == 1).flatten(1).sum(1)
    assert bool((nb == nw).all()) and int(nb.max()) > 3 and int(nb.min()) >= 0           # equal stones -> BLACK to move
    occ = bd != 2
    rows, cols = occ.any(2).sum(1), occ.any(1).sum(1)
    assert int(rows.max()) <= N - 1 and int(cols.max()) <= N - 1                           # inside a rectangle of side <= N-2(+)
    venv = HexVecEnv(board_size=N, num_envs=G, variant="selfplay", seed=3, output="torch", sample_board=True)
    obs = venv.reset()
    stones0 = (obs != 0).flatten(1).sum(1)
    assert int(stones0.max()) > 2 and bool(((obs == 0).flatten(1) == venv.action_masks()).all())
    agent_white = venv.batch.export_state()["agent"] == 1
    own, opp = (obs == -1).flatten(1).sum(1), (obs == 1).flatten(1).sum(1)
    assert bool((opp - own == agent_white.long()).all())       # the opponent opened exactly where the agent plays WHITE
    total_done = 0
    for t in range(40):
        masks = venv.action_masks()
        acts = torch.argmax(masks.to(torch.uint8), dim=1).to(torch.int32)
        obs, rew, done, infos = venv.step(acts)
        total_done += int(done.sum())
        assert bool(((obs == 0).flatten(1) == venv.action_masks()).all())
        own, opp = (obs == -1).flatten(1).sum(1), (obs == 1).flatten(1).sum(1)
        assert bool((opp - own == agent_white.long()).all())   # always the agent's turn, on consistent positions
        assert bool((rew[~done] == 0).all()) and bool((rew[done].abs() == 1).all())
    assert total_done > G // 4 and venv.episode_stats()["invalid_ends"] == 0


@pytest.mark.parametrize("N", list(range(3, 20)))
def test_every_board_size_instantiation(make, N):
    """Every template instantiation (N = 3..19), both variants, short fused rollouts against the oracle."""
    parity.versus_oracle(make, hexref.KIND_SELFPLAY_B, N, 97, min(N * N // 2 + 4, 40), seed=100 + N, fused=True, agent_mode=2,
                         check_state_every=7)
    parity.versus_oracle(make, hexref.KIND_ENV_A, N, 65, min(N * N // 2 + 4, 30), seed=200 + N, fused=True, opponent_first=bool(N % 2),
                         check_state_every=7)


def test_misaligned_output_buffers():
    """Caller buffers that are not 16-byte aligned (slices of a larger buffer) take the byte-wise store path: same bytes."""
    import torch
    from hex_gym_env_b200 import HexBatch, VARIANT_B
    N, G = 11, 333
    a = HexBatch(N, G, variant=VARIANT_B, device=0, seed=2, agent_mode=2)
    b = HexBatch(N, G, variant=VARIANT_B, device=0, seed=2, agent_mode=2)
    big_o = torch.zeros(G * N * N + 7, dtype=torch.int8, device="cuda")
    big_m = torch.zeros(G * N * N + 7, dtype=torch.uint8, device="cuda")
    obs_u = big_o[3:3 + G * N * N].view(G, N, N)
    mask_u = big_m[5:5 + G * N * N].view(G, N * N)
    assert obs_u.data_ptr() % 16 != 0 and mask_u.data_ptr() % 16 != 0
    a.reset(); b.reset(obs=obs_u, mask=mask_u)
    for t in range(40):
        o = a.step()
        b.step(obs=obs_u, mask=mask_u)
        assert torch.equal(o["obs"], obs_u) and torch.equal(o["mask"], mask_u), t
    assert int(big_o[:3].abs().sum()) == 0 and int(big_o[3 + G * N * N:].abs().sum()) == 0      # nothing written outside the slice
    assert int(big_m[:5].sum()) == 0 and int(big_m[5 + G * N * N:].sum()) == 0


def test_capture_steps_graph():
    """HexBatch.capture_steps: K steps replayed from a CUDA graph == K eager steps (fused agent and external actions)."""
    import torch
    from hex_gym_env_b200 import HexBatch
    N, G, K = 7, 3000, 8
    a = HexBatch(N, G, variant=1, device=0, seed=12, agent_mode=2)
    b = HexBatch(N, G, variant=1, device=0, seed=12, agent_mode=2)
    a.reset(); b.reset()
    g = a.capture_steps(K)
    g.replay(); g.replay()
    torch.cuda.synchronize()
    for _ in range(1 + 2 * K):
        o = b.step()
    for k in ("obs", "mask", "reward", "done"):
        assert torch.equal(g.outputs[k], o[k]), k
    assert torch.equal(a.stats(), b.stats())
    acts = torch.zeros(G, dtype=torch.int32, device="cuda")
    g1 = a.capture_steps(1, actions=acts)          # the eager step inside plays cell 0 everywhere
    b.step(acts)
    for t in range(5):
        acts.copy_(a.sample_actions(torch.full((G,), 0.37, dtype=torch.float64, device="cuda")))
        g1.replay()
        o = b.step(acts)
        for k in ("obs", "mask", "reward", "done"):
            assert torch.equal(g1.outputs[k], o[k]), (k, t)


@pytest.mark.parametrize("seed", range(48))
def test_api_fuzz(make, seed):
    """Random walks over the step / reset API (fused, external and illegal moves, injected draws, masked resets)."""
    parity.api_fuzz(make, seed, T=60)


@pytest.mark.parametrize("N,variant_a", [(3, False), (4, True), (5, True), (6, False), (7, True), (8, False), (9, True), (10, False), (11, False),
                                         (12, True), (13, False), (14, False), (15, True), (16, False), (17, False), (18, True), (19, False)])
def test_sampler_and_views(make, N, variant_a):
    parity.sampler_and_views(make, N, variant_a, seed=N)


@pytest.mark.parametrize("N,G,T", [(11, 2048, 1500), (19, 256, 700), (5, 4096, 1200)])
def test_long_run_vs_oracle(make, N, G, T):
    """Dozens of episodes per game (auto-reset): reward / done / actions every step, the whole exported state every 100 steps
    and the episode statistics at the end, bit-exact against the oracle."""
    ref = hexref.RefBatch(hexref.KIND_SELFPLAY_B, N, G, seed=77, agent_mode=2)
    env = make(hexref.KIND_SELFPLAY_B, N, G, seed=77, agent_mode=2)
    ref.reset(); env.reset()
    for t in range(T):
        r, o = ref.step(), env.step()
        for k in ("reward", "done", "actions"):
            parity.eq(o[k], r[k], "long run N=%d t=%d %s" % (N, t, k))
        if t % 100 == 99:
            parity.eq(o["obs"], r["obs"], "long run obs t=%d" % t)
            parity.eq(o["mask"], r["mask"], "long run mask t=%d" % t)
            re_, e = ref.export(), env.export()
            for k in parity.STATE_KEYS:
                parity.eq(e[k], re_[k], "long run N=%d t=%d %s" % (N, t, k))
    parity.eq(env.stats(), ref.stats(), "long run stats")
    assert ref.stats()[0] > 3 * G


def test_config1_ten_thousand_raw_games(make):
    """BASELINE config 1: 10,000 raw 5x5 HexGame instances, random-vs-random, state compared after every ply."""
    parity.raw_random_games(make, 5, 10000, seed=1)


@pytest.mark.parametrize("N", [3, 7, 11, 14])
def test_raw_random_games(make, N):
    parity.raw_random_games(make, N, 600, seed=N)


@pytest.mark.parametrize("name", golden_files("saturation_"))
def test_golden_label_saturation(make, name):
    """~110 distinct region labels per colour on 18x18 / 19x19 (the packed state holds 7-bit labels), then everything merges:
    hexb_ply vs the unmodified reference's snapshots and vs the oracle after every ply."""
    parity.golden_saturation(make, name)


@pytest.mark.parametrize("name", golden_files("presetreset_"))
def test_golden_preset_resets(name):
    """hexb_import_labels: HexGame.__init__ with connected_stones (the cached planes of HexEnv.reset's later calls)."""
    from gpu_adapter import GpuBatch
    parity.golden_preset_resets(lambda kind, N, G: GpuBatch(0 if kind == hexref.KIND_GAME_A else 1, N, G, raw=True), name)


@pytest.mark.parametrize("name", golden_files("oppredict_"))
def test_opponent_predict_batched(make, name):
    """hexb_set_opponent_eps + hexb_half_step: variant-A HexEnv(opponent_policy="opponent_predict", eps=...) for a batch against the
    unmodified reference run one env per game."""
    parity.golden_oppredict_batched(make, name)


@pytest.mark.parametrize("kind,N,seed", [(hexref.KIND_SELFPLAY_B, 5, 12), (hexref.KIND_SELFPLAY_B, 11, 13), (hexref.KIND_ENV_A, 4, 14),
                                         (hexref.KIND_ENV_A, 7, 15)])
def test_opponent_modes_fuzz(make, kind, N, seed):
    """set_eval / the eps of opponent_predict switched at random moments of a random split-step run, GPU vs oracle."""
    parity.opponent_modes_fuzz(make, kind, N, 600, 60, seed)
