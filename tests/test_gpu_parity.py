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
