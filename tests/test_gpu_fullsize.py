"""GPU (-m gpu): the BASELINE.json configurations at their FULL sizes, through the C ABI.

The oracle cannot follow millions of games, so each test combines
  * bit-exact comparison with the oracle on WINDOWS of the global batch (the oracle simulates global game indices
    [start, start+len) with the same seed: per-game Philox streams are keyed by the global index), placed on chunk, tile and
    shard boundaries and at the very end of the batch,
  * size-independent properties over ALL games (tests/properties.py, itself pinned against the oracle on the CPU):
    mask == empty cells, stone counts vs the agent's colour, done <=> reward != 0, nobody connected in a live position, the
    winner connected in every terminal observation (brute-force dilation, independent of the label bookkeeping),
    label-plane consistency of the exported state, and conservation of the episode statistics,
  * sharding invariance at full size: 8 shards (the 8-GPU partition) == the unsharded batch, byte for byte.
"""
import numpy as np
import pytest
import torch

import properties
from oracle import hexref

pytestmark = pytest.mark.gpu


def _windows(G, shard):
    """(start, length) windows: the first chunk, a ragged span across a 32-game chunk and a 128-game tile boundary, a span
    across a per-GPU shard boundary, the middle, and the last games of the batch."""
    return [(0, 32), (32 * 5 - 7, 45), (128 * 1000 - 3, 9), (shard - 16, 32), (G // 2 + 11, 33), (G - 40, 40)]


def _run_config(N, G, T, variant, kind, agent_mode, shard, check_every, label_check):
    from hex_gym_env_b200 import HexBatch
    env = HexBatch(N, G, variant=variant, device=0, seed=0, agent_mode=agent_mode, auto_reset=True)
    wins = _windows(G, shard)
    refs = [hexref.RefBatch(kind, N, n, seed=0, game_offset=s, agent_mode=agent_mode) for s, n in wins]
    obs, mask = env.reset()
    for (s, n), ref in zip(wins, refs):
        ro, rm = ref.reset()
        assert np.array_equal(obs[s:s + n].cpu().numpy(), ro) and np.array_equal(mask[s:s + n].cpu().numpy(), rm), ("reset", s)
    agent = env.export_state()["agent"] if variant == 1 else None
    term = torch.zeros((G, N, N), dtype=torch.int8, device="cuda")
    n_done = 0
    for t in range(T):
        o = env.step(term_obs=term, want_actions=True)
        for (s, n), ref in zip(wins, refs):
            r = ref.step(want_term=True)
            for k in ("obs", "mask", "reward", "done", "actions"):
                assert np.array_equal(o[k][s:s + n].cpu().numpy(), r[k]), "%s differs from the oracle in window %d at step %d" % (k, s, t)
            d = r["done"].astype(bool)
            assert np.array_equal(term[s:s + n].cpu().numpy()[d], r["term_obs"][d]), ("term_obs", s, t)
        n_done += int(o["done"].sum().item())
        if variant == 1 and (t % check_every == check_every - 1 or t == T - 1):
            properties.check_selfplay_step(o["obs"], o["mask"], o["reward"], o["done"], agent, term)
    st = dict(zip(("episodes", "black_wins", "white_wins", "agent_wins", "episode_plies", "invalid_ends", "env_steps", "plies"),
                  env.stats().cpu().tolist()))
    assert st["env_steps"] == G * T and st["episodes"] == n_done and st["invalid_ends"] == 0
    assert st["black_wins"] + st["white_wins"] == st["episodes"]
    assert st["episodes"] < 1000 or 0.4 < st["agent_wins"] / st["episodes"] < 0.6
    assert G * T <= st["plies"] <= 2 * G * T + G
    if label_check:
        e = env.export_state()
        properties.check_exported_labels(e["board"], e["regions"], 0 if variant == 1 else 2)
        for (s, n), ref in zip(wins, refs):
            re_ = ref.export()
            for k in ("board", "regions", "region_counter", "cur", "done", "winner", "agent"):
                assert np.array_equal(e[k][s:s + n].cpu().numpy(), re_[k]), ("state", k, s)
        del e
    for (s, n), ref in zip(wins[:1], refs[:1]):
        assert ref.stats()[6] == n * T
    env.close()
    return st


def test_config3_full_size():
    """BASELINE config 3: 11x11 SelfPlayEnv, random opponent, agent colour random per game, 1,048,576 games, 150 steps
    (almost three mean episodes of 54 env steps, so most games restarted twice)."""
    st = _run_config(11, 1 << 20, 150, 1, hexref.KIND_SELFPLAY_B, 2, (1 << 20) // 8, check_every=25, label_check=True)
    assert st["episodes"] > 2 * (1 << 20)


def test_config5_full_size():
    """BASELINE config 5: 19x19 SelfPlayEnv, 4,194,304 games."""
    st = _run_config(19, 1 << 22, 36, 1, hexref.KIND_SELFPLAY_B, 2, (1 << 22) // 8, check_every=18, label_check=False)
    assert st["episodes"] < (1 << 22) // 100                   # 36 env steps = 72 plies: hardly any game has ended


def test_config5_runs_to_the_end_of_the_games():
    """19x19 at 524,288 games for 200 steps (> the 167-step mean episode): every path of the long games is exercised, labels
    grow past 64, and the exported label planes stay consistent."""
    st = _run_config(19, 1 << 19, 200, 1, hexref.KIND_SELFPLAY_B, 2, (1 << 19) // 8, check_every=100, label_check=True)
    assert st["episodes"] > (1 << 19) // 2


def test_config2_full_size():
    """BASELINE config 2: 7x7 variant-A HexEnv + random_policy opponent with action masks, 65,536 games, 2,000 steps."""
    from hex_gym_env_b200 import HexBatch
    N, G, T = 7, 65536, 2000
    env = HexBatch(N, G, variant=0, device=0, seed=0, auto_reset=True)
    wins = [(0, 32), (4096 - 5, 10), (G - 33, 33)]
    refs = [hexref.RefBatch(hexref.KIND_ENV_A, N, n, seed=0, game_offset=s) for s, n in wins]
    env.reset()
    for ref in refs:
        ref.reset()
    n_done = 0
    for t in range(T):
        o = env.step(want_actions=True)
        if t % 10 == 0 or t > T - 5:
            host = {k: o[k].cpu().numpy() for k in ("obs", "mask", "reward", "done", "actions")}
        for (s, n), ref in zip(wins, refs):
            r = ref.step()
            if t % 10 == 0 or t > T - 5:
                for k in host:
                    assert np.array_equal(host[k][s:s + n], r[k]), (k, s, t)
        n_done += int(o["done"].sum().item()) if t % 100 == 0 else 0
        if t % 250 == 0:
            G_, obs = G, o["obs"]
            assert torch.equal(o["mask"].view(G_, N, N) != 0, obs == 2)          # get_action_mask, HexGame.py:203-204
            nb, nw = (obs == 0).sum(dim=(1, 2)), (obs == 1).sum(dim=(1, 2))
            assert torch.equal(nb, nw)                                           # agent BLACK to move: equal stone counts
            assert not properties.connects(obs == 0, 0).any() and not properties.connects(obs == 1, 1).any()
    st = env.stats().cpu().tolist()
    assert st[6] == G * T and st[1] + st[2] == st[0] and st[5] == 0
    e = env.export_state()
    properties.check_exported_labels(e["board"], e["regions"], 2)
    for (s, n), ref in zip(wins, refs):
        re_ = ref.export()
        for k in ("board", "regions", "region_counter", "cur", "done", "winner"):
            assert np.array_equal(e[k][s:s + n].cpu().numpy(), re_[k]), ("state", k, s)


def test_sharding_invariance_full_size():
    """Config 3 as the 8-GPU partition lays it out (8 shards of 131,072 games, here on one device) == the unsharded
    1,048,576-game batch: every output byte of every step, the packed game state, and the summed statistics."""
    from hex_gym_env_b200 import HexBatch
    N, G, R, T = 11, 1 << 20, 8, 40
    S = G // R
    whole = HexBatch(N, G, variant=1, device=0, seed=3, agent_mode=2)
    shards = [HexBatch(N, S, variant=1, device=0, seed=3, agent_mode=2, game_offset=r * S) for r in range(R)]
    whole.reset()
    for sh in shards:
        sh.reset()
    for t in range(T):
        w = whole.step(want_actions=True)
        for r, sh in enumerate(shards):
            o = sh.step(want_actions=True)
            for k in ("obs", "mask", "reward", "done", "actions"):
                assert torch.equal(w[k][r * S:(r + 1) * S], o[k]), (k, r, t)
    total = sum(sh.stats() for sh in shards)
    assert torch.equal(total, whole.stats())
    # the packed state is chunk-major (32 games per block), so shard r's game blocks are a contiguous slice of the whole's
    ws = whole.state_dict()["state"]
    C, W = N * N, (N * N + 31) // 32
    per_shard = S // 32 * 32 * (C + 4 * (W + 2))                                 # labels + records per chunk
    for r, sh in enumerate(shards):
        assert torch.equal(sh.state_dict()["state"][:per_shard], ws[r * per_shard:(r + 1) * per_shard]), r
    assert properties.checksum(torch.cat([sh._out["obs"] for sh in shards])) == properties.checksum(whole._out["obs"])


def test_fullsize_host_transports_and_float32_observations():
    """1 Mi games of 11x11 (BASELINE config 3's size): the host-buffer step with every transport (plain DMA, all packed, the
    library's own search) and the float32 observation dtype return the device step's bytes, whole arrays compared."""
    from hex_gym_env_b200 import HexBatch
    N, G, T = 11, 1 << 20, 26
    dev = HexBatch(N, G, variant=1, device=0, seed=3, agent_mode=2)
    f32 = HexBatch(N, G, variant=1, device=0, seed=3, agent_mode=2, obs_dtype=torch.float32)
    hosts = []
    for frac in (1.0, 0.0, -1.0):
        h = HexBatch(N, G, variant=1, device=0, seed=3, agent_mode=2)
        h.set_host_transport(frac)
        h.reset()
        hosts.append((frac, h, h.pinned_io()))
    dev.reset(); f32.reset()
    for t in range(T):
        o = dev.step()
        of = f32.step()
        check = t in (0, 1, 12, T - 1)      # (the adaptive handle walks through all its candidate splits over these 26 calls)
        if check:
            assert torch.equal(of["obs"], o["obs"].float()) and torch.equal(of["mask"], o["mask"]), ("f32", t)
            want = {k: o[k].cpu().numpy() for k in ("obs", "mask", "reward", "done")}
        for frac, h, io in hosts:
            if check:
                io["obs"].fill_(77); io["mask"].fill_(77)
            h.step_host(None, io)
            if check:
                for k in ("obs", "mask", "reward", "done"):
                    assert np.array_equal(io[k].numpy(), want[k]), (k, t, frac)
    assert 0.0 <= hosts[2][1].host_transport() <= 1.0
    for _, h, _ in hosts:
        assert torch.equal(h.stats(), dev.stats())
