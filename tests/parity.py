"""Shared parity drivers: run an implementation of the batched simulator next to the golden vectors and the oracle.

`make(kind, N, G, **kw)` must return an object with the RefBatch method surface (reset / step / ply / export / stats,
numpy in and out). tests/test_emu_parity.py passes the host emulator of the kernel phases (CPU); tests/test_gpu_parity.py
passes the CUDA library through its C ABI (GPU). Every comparison is bit-exact (integer / byte work; rewards are small
integers in f32)."""
import os

import numpy as np

from conftest import GOLDEN
from oracle import hexref

STATE_KEYS = ("board", "regions", "region_counter", "cur", "done", "winner", "agent", "draws")


def eq(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape or not np.array_equal(a, b):
        bad = np.argwhere(a != b) if a.shape == b.shape else None
        raise AssertionError("%s differs%s" % (what, "" if bad is None else " at %s (first of %d)" % (bad[0], len(bad))))


def golden_raw_game(make, name):
    z = np.load(os.path.join(GOLDEN, name))
    N = int(z["N"])
    moves = z["moves"]
    n_games, T = moves.shape
    env = make(hexref.KIND_GAME_A, N, n_games)
    env.reset()
    for t in range(T):
        ret = env.ply(moves[:, t])
        e = env.export()
        eq(ret, z["ret"][:, t], "%s ret t=%d" % (name, t))
        eq(e["board"], z["board"][:, t].astype(np.float64), "%s board t=%d" % (name, t))
        eq(e["regions"], z["regions"][:, t].astype(np.float64), "%s regions t=%d" % (name, t))
        eq(e["region_counter"], z["counter"][:, t].astype(np.float64), "%s counter t=%d" % (name, t))
        eq(e["cur"], z["cur"][:, t], "%s cur t=%d" % (name, t))
        eq(e["done"], z["done"][:, t], "%s done t=%d" % (name, t))
        eq(e["winner"], z["winner"][:, t], "%s winner t=%d" % (name, t))


def golden_rollout(make, name):
    z = np.load(os.path.join(GOLDEN, name))
    N, seed, fused = int(z["N"]), int(z["seed"]), int(z["fused"])
    T, G = z["actions"].shape
    if os.path.basename(name).startswith("selfplay"):    # a fixture's file name, or the path of a trace written elsewhere
        env = make(hexref.KIND_SELFPLAY_B, N, G, seed=seed, agent_mode=int(z["agent_mode"]))
    else:
        env = make(hexref.KIND_ENV_A, N, G, seed=seed, opponent_first=bool(int(z["opponent_first"])))
    obs0, mask0 = env.reset()
    eq(obs0, z["obs0"], name + " obs0")
    eq(mask0, z["mask0"], name + " mask0")
    e = env.export()
    eq(e["agent"], z["agent"], name + " agent")
    eq(e["draws"], z["draws0"], name + " draws0")
    for t in range(T):
        o = env.step(None if fused else z["actions"][t], want_term=True)
        w = "%s t=%d " % (name, t)
        eq(o["actions"], z["actions"][t], w + "actions")
        eq(o["reward"], z["reward"][t], w + "reward")
        eq(o["done"], z["done"][t], w + "done")
        eq(o["obs"], z["obs"][t], w + "obs")
        eq(o["mask"], z["mask"][t], w + "mask")
        d = z["done"][t].astype(bool)
        eq(o["term_obs"][d], z["term_obs"][t][d], w + "term_obs")
        e = env.export()
        eq(e["regions"], z["regions"][t].astype(np.float64), w + "regions")
        eq(e["region_counter"], z["counter"][t].astype(np.float64), w + "counter")
        eq(e["cur"], z["sim_cur"][t], w + "cur")
        eq(e["draws"], z["draws"][t], w + "draws")


def versus_oracle(make, kind, N, G, T, seed=0, game_offset=0, fused=True, auto_reset=True, illegal_rate=0.03,
                  check_state_every=1, **kw):
    """Side-by-side rollout of `make(...)` and the oracle on the same seeded inputs."""
    ref = hexref.RefBatch(kind, N, G, seed=seed, game_offset=game_offset, **kw)
    env = make(kind, N, G, seed=seed, game_offset=game_offset, auto_reset=auto_reset, **kw)
    rs = np.random.RandomState(seed + 17)
    if hasattr(env, "enable_info"):
        env.enable_info()
    ro, rm = ref.reset()
    o, m = env.reset()
    eq(o, ro, "reset obs")
    eq(m, rm, "reset mask")
    C = N * N
    for t in range(T):
        if fused:
            acts = None
        else:
            # uniformly random legal action from the oracle's mask, sometimes an arbitrary (possibly illegal) cell
            u = rs.rand(G)
            cnt = rm.sum(1).astype(np.int64)
            k = np.minimum((u * cnt).astype(np.int64), np.maximum(cnt - 1, 0))
            order = np.argsort(-rm.astype(np.int8), axis=1, kind="stable")
            acts = order[np.arange(G), k].astype(np.int32)
            bad = rs.rand(G) < illegal_rate
            acts[bad] = rs.randint(-2, C + 2, size=int(bad.sum()))
        r = ref.step(acts, auto_reset=auto_reset, want_term=True)
        o = env.step(acts, want_term=True)
        w = "N=%d t=%d " % (N, t)
        for key in ("actions", "reward", "done", "obs", "mask"):
            if key == "actions" and not fused:
                continue
            eq(o[key], r[key], w + key)
        if hasattr(env, "enable_info"):
            (eo, ew), (fo, fw) = env.info(), ref.info()
            eq(eo, fo, w + "info last_move_opponent")
            eq(ew, fw, w + "info winner")
        d = r["done"].astype(bool)
        # with auto-reset the terminal observation comes through term_obs; without it, it is the observation itself
        eq(o["term_obs"][d] if auto_reset else o["obs"][d], r["term_obs"][d], w + "term_obs")
        rm = r["mask"]
        if check_state_every and t % check_state_every == 0:
            re_, e = ref.export(), env.export()
            for key in STATE_KEYS:
                eq(e[key], re_[key], w + key)
    eq(env.stats(), ref.stats(), "stats")
    return ref, env


def snake_moves(N):
    """Deterministic adversarial game (BASELINE config 5): BLACK builds ONE serpentine chain through every even row (joined at
    alternating ends of the odd rows), first placing every other stone of the path (about N*N/4 isolated stones = as many
    distinct labels as the board can hold), then filling the gaps so that every move merges two groups and the final moves
    relabel a chain of ~N*N/2 stones; WHITE fills the rest of the odd rows. Worst case for a flood-fill based checker."""
    path = []
    for r in range(N):
        if r % 2 == 0:
            xs = range(N) if (r // 2) % 2 == 0 else range(N - 1, -1, -1)
            path += [r * N + x for x in xs]
        else:
            path.append(r * N + (N - 1 if (r // 2) % 2 == 0 else 0))
    black = path[0::2] + path[1::2]
    taken = set(path)
    white = [c for c in range(N * N) if c not in taken]
    moves = []
    for i in range(min(len(black), len(white))):
        moves += [black[i], white[i]]
    return moves


def snake_chain(make, N):
    moves = snake_moves(N)
    G = 40
    env = make(hexref.KIND_GAME_A, N, G)
    ref = hexref.RefBatch(hexref.KIND_GAME_A, N, G)
    env.reset()
    rs = np.random.RandomState(1)
    lag = rs.randint(0, 3, size=G)          # games run the same list with small offsets so that lanes are out of step
    for t in range(len(moves) + 2):
        idx = np.clip(t - lag, 0, len(moves) - 1)
        a = np.array(moves, np.int32)[idx]
        eq(env.ply(a), ref.ply(a), "snake ret t=%d" % t)
        if t % 7 == 0 or t >= len(moves) - 3:
            e, r = env.export(), ref.export()
            for k in ("board", "regions", "region_counter", "cur", "done", "winner"):
                eq(e[k], r[k], "snake %s t=%d" % (k, t))
    assert ref.export()["region_counter"].max() >= N * N // 5   # the label range really was exercised


def golden_oppmodel(make, name):
    """Caller-driven opponent (hexb_half_step) against the reference run with OpponentPolicy opponents (scripted models).
    The opponent's actions are re-derived here from the implementation's OWN side-to-move observation with the same scripted
    rule the reference's models used, so the opponent view, the 80/20 opponent choice and the half steps are all pinned.
    evalpool_* fixtures also switch SelfPlayEnv.set_eval on and off in mid-run (eval_at[t]: the call right before step t): the
    evaluation cycle through the pool (SelfplayWrapper.py:92-96) must hand every episode the same pool entry."""
    from oracle.scripted import scripted_choice
    z = np.load(os.path.join(GOLDEN, name))
    N, seed, am, pool = int(z["N"]), int(z["seed"]), int(z["agent_mode"]), int(z["pool"])
    T, G = z["actions"].shape
    eval_at = z["eval_at"] if "eval_at" in z.files else None
    env = make(hexref.KIND_SELFPLAY_B, N, G, seed=seed, agent_mode=am, manual_opponent=True, pool_size=pool)
    env.reset()

    def opponent_pass(expect_a, expect_model, what):
        tm, idx = env.opp_state()
        eq(tm == 1, expect_a >= 0, what + " who waits for the opponent")
        obs1, mask1 = env.view1()
        acts = np.zeros(G, np.int32)
        for g in np.flatnonzero(tm == 1):
            acts[g] = scripted_choice(obs1[g], mask1[g])
            assert acts[g] == expect_a[g], (what, g, acts[g], expect_a[g])
            assert idx[g] == expect_model[g], (what, "opponent index", g, idx[g], expect_model[g])
        return env.half_step(1, acts, want_term=True)

    opponent_pass(z["opp0_action"], z["opp0_model"], name + " opening")
    obs, mask = env.view1()
    eq(obs, z["obs0"], name + " obs0")
    eq(mask, z["mask0"], name + " mask0")
    e = env.export()
    eq(e["agent"], z["agent"], name + " agent")
    eq(e["draws"], z["draws0"], name + " draws0")
    for t in range(T):
        w = "%s t=%d " % (name, t)
        if eval_at is not None and eval_at[t] >= 0:
            env.set_eval(bool(eval_at[t]))
        h = env.half_step(0, z["actions"][t], want_term=True)
        reward, done, term = h["reward"].copy(), h["done"].astype(bool), h["term_obs"].copy()
        for j in (0, 1):
            h = opponent_pass(z["opp_actions"][t, :, j], z["opp_model"][t, :, j], w + "pass %d" % j)
            reward += h["reward"]
            d = h["done"].astype(bool)
            term[d] = h["term_obs"][d]
            done |= d
        tm, _ = env.opp_state()
        assert (tm == 0).all(), w + "every game back at the agent"
        eq(reward, z["reward"][t], w + "reward")
        eq(done, z["done"][t].astype(bool), w + "done")
        eq(term[done], z["term_obs"][t][done], w + "term_obs")
        obs, mask = env.view1()
        eq(obs, z["obs"][t], w + "obs")
        eq(mask, z["mask"][t], w + "mask")
        e = env.export()
        eq(e["regions"], z["regions"][t].astype(np.float64), w + "regions")
        eq(e["region_counter"], z["counter"][t].astype(np.float64), w + "counter")
        eq(e["cur"], z["sim_cur"][t], w + "cur")
        eq(e["draws"], z["draws"][t], w + "draws")


def rollout_equals_steps(make, kind, N, G, T, seed=0, **kw):
    """hexb_rollout(T) is bit-identical to T calls of hexb_step(actions=None), and both match the oracle."""
    a = make(kind, N, G, seed=seed, **kw)
    b = make(kind, N, G, seed=seed, **kw)
    ref = hexref.RefBatch(kind, N, G, seed=seed, **kw)
    a.reset(); b.reset(); ref.reset()
    for rep in range(2):    # two launches back to back: the state carried between launches is right too
        ro = a.rollout(T, want_term=True)
        for t in range(T):
            so = b.step(want_term=True)
            r = ref.step(want_term=True)
            w = "rollout rep=%d t=%d " % (rep, t)
            for key, okey in (("obs", "obs"), ("mask", "mask"), ("reward", "reward"), ("done", "done"), ("actions", "actions")):
                eq(ro[okey][t], so[key], w + key + " vs steps")
                eq(ro[okey][t], r[key], w + key + " vs oracle")
            d = r["done"].astype(bool)
            eq(ro["term_obs"][t][d], r["term_obs"][d], w + "term_obs")
        ea, eb = a.export(), b.export()
        for key in STATE_KEYS:
            eq(ea[key], eb[key], "rollout state " + key)
        eq(a.stats(), b.stats(), "rollout stats")


def random_start_boards(rs, G, N, variant_b=True):
    """Random positions with equally many stones of both colours (BLACK to move), like HexEnv.random_board produces."""
    true_codes = np.full((G, N * N), 2, np.int8)
    for g in range(G):
        k = int(rs.randint(0, N * N // 3)) * 2
        cells = rs.permutation(N * N)[:k]
        true_codes[g, cells[:k // 2]] = 0
        true_codes[g, cells[k // 2:]] = 1
    true_codes = true_codes.reshape(G, N, N)
    own = np.where(true_codes == 0, -1, np.where(true_codes == 1, 1, 0)).astype(np.int8) if variant_b else true_codes
    return true_codes, own


def sample_board_flow(make, N, G, T, seed=0, agent_mode=2):
    """SelfPlayEnv with sample_board=True and the random opponent, as a split-step flow: agent ply, opponent reply, finished
    games restart from random positions (import) and the opponent catches up where it is to move. Implementation vs oracle."""
    env = make(hexref.KIND_SELFPLAY_B, N, G, seed=seed, agent_mode=agent_mode, manual_opponent=True)
    ref = hexref.RefBatch(hexref.KIND_SELFPLAY_B, N, G, seed=seed, agent_mode=agent_mode, manual_opponent=True)
    rs = np.random.RandomState(seed + 5)
    env.reset(); ref.reset()
    tc, own = random_start_boards(rs, G, N)
    env.import_boards(tc); ref.env_set_board(own)
    env.half_step(1, None); ref.half_step(1, None)
    for t in range(T):
        obs, mask = env.view1()
        robs, rmask = ref.view1()
        eq(obs, robs, "sample_board obs t=%d" % t)
        eq(mask, rmask, "sample_board mask t=%d" % t)
        cnt = np.maximum(mask.sum(1), 1)
        k = (rs.rand(G) * cnt).astype(np.int64)
        acts = np.argsort(-mask.astype(np.int8), axis=1, kind="stable")[np.arange(G), k].astype(np.int32)
        done = np.zeros(G, bool)
        for side, a in ((0, acts), (1, None)):
            o, r = env.half_step(side, a), ref.half_step(side, a)
            for key in ("reward", "done", "to_move"):
                eq(o[key], r[key], "sample_board %s t=%d side=%d" % (key, t, side))
            done |= r["done"].astype(bool)
        tc, own = random_start_boards(rs, G, N)
        env.import_boards(tc, import_mask=done.astype(np.uint8)); ref.env_set_board(own, done.astype(np.uint8))
        o, r = env.half_step(1, None), ref.half_step(1, None)
        eq(o["to_move"], r["to_move"], "sample_board to_move after catch-up t=%d" % t)
        if t % 4 == 0:
            e, re_ = env.export(), ref.export()
            for key in ("regions", "region_counter", "cur", "done", "winner", "agent", "draws"):
                eq(e[key], re_[key], "sample_board %s t=%d" % (key, t))
    eq(env.stats(), ref.stats(), "sample_board stats")


def golden_preset(make_raw, name):
    """Preset-board construction (raster-order label rebuild) against the reference's own HexGame.__init__."""
    z = np.load(os.path.join(GOLDEN, name))
    N = int(z["N"])
    tc = z["board_true"]
    for variant, kind in (("A", hexref.KIND_GAME_A), ("B", hexref.KIND_GAME_B)):
        env = make_raw(kind, N, tc.shape[0])
        e = env.export()
        eq(e["regions"], z["regions_" + variant].astype(np.float64), "%s %s regions" % (name, variant))
        eq(e["region_counter"], z["counter_" + variant].astype(np.float64), "%s %s counter" % (name, variant))


def api_fuzz(make, seed, T=50):
    """Random walk over the step / reset API against the oracle: fused and external (sometimes illegal) agent moves, injected
    opponent draws (opp_u), masked resets with and without injected opening draws (open_u), with and without auto-reset."""
    rs = np.random.RandomState(1000 + seed)
    N = int(rs.choice([3, 4, 5, 6, 7, 8, 9, 11, 13]))
    G = int(rs.choice([1, 31, 32, 33, 97, 128, 200]))
    variant_a = bool(rs.rand() < 0.35)
    auto_reset = bool(rs.rand() < 0.6)
    kw = dict(opponent_first=bool(rs.rand() < 0.5)) if variant_a else dict(agent_mode=int(rs.randint(0, 3)))
    kind = hexref.KIND_ENV_A if variant_a else hexref.KIND_SELFPLAY_B
    off = int(rs.randint(0, 1 << 40))
    ref = hexref.RefBatch(kind, N, G, seed=seed, game_offset=off, **kw)
    env = make(kind, N, G, seed=seed, game_offset=off, auto_reset=auto_reset, **kw)
    what = "fuzz seed=%d N=%d G=%d %s auto_reset=%d %r" % (seed, N, G, "A" if variant_a else "B", auto_reset, kw)
    ro, rm = ref.reset()
    o, m = env.reset()
    eq(o, ro, what + " reset obs")
    eq(m, rm, what + " reset mask")
    C = N * N
    for t in range(T):
        op = rs.choice(["fused", "actions", "actions_u", "reset_mask", "reset_mask_u"], p=[0.3, 0.3, 0.2, 0.1, 0.1])
        w = "%s t=%d op=%s " % (what, t, op)
        if op.startswith("reset"):
            mask = (rs.rand(G) < 0.3).astype(np.uint8)
            ou = rs.rand(G) if op.endswith("_u") else None
            ro, rm = ref.reset(mask, ou)
            o, m = env.reset(mask, ou)
            eq(o, ro, w + "obs")
            eq(m, rm, w + "mask")
        else:
            acts, u = None, None
            if op != "fused":
                cnt = rm.sum(1).astype(np.int64)
                k = np.minimum((rs.rand(G) * cnt).astype(np.int64), np.maximum(cnt - 1, 0))
                acts = np.argsort(-rm.astype(np.int8), axis=1, kind="stable")[np.arange(G), k].astype(np.int32)
                bad = rs.rand(G) < 0.05
                acts[bad] = rs.randint(-2, C + 2, size=int(bad.sum()))
            if op == "actions_u":
                u = rs.rand(G, 2)
                u[rs.rand(G) < 0.1, 0] = 0.0
                u[rs.rand(G) < 0.1, 0] = np.nextafter(1.0, 0.0)
            r = ref.step(acts, u, auto_reset=auto_reset, want_term=True)
            e = env.step(acts, u, want_term=True)
            for key in ("reward", "done", "obs", "mask"):
                eq(e[key], r[key], w + key)
            if acts is None:
                eq(e["actions"], r["actions"], w + "actions")
            d = r["done"].astype(bool)
            if auto_reset:
                eq(e["term_obs"][d], r["term_obs"][d], w + "term_obs")
            rm = r["mask"]
        if t % 5 == 4:
            re_, ee = ref.export(), env.export()
            for key in STATE_KEYS:
                eq(ee[key], re_[key], w + key)
    eq(env.stats(), ref.stats(), what + " stats")


def sampler_and_views(make, N, variant_a, seed=0):
    """hexb_sample_actions and hexb_encode on their own, at positions reached by random play: the sampled action is the
    int(u * n_empty)-th legal cell of the requested view in row-major order (BaseRandomPolicy.choose_action, SelfplayWrapper.py:17-22;
    random_policy, minihex/__init__.py:8-12), for view 0 (the agent's) and view 1 (the side to move's); and encode(view 0) is
    what the last step returned."""
    G = 150
    kind = hexref.KIND_ENV_A if variant_a else hexref.KIND_SELFPLAY_B
    kw = {} if variant_a else dict(agent_mode=2)
    env = make(kind, N, G, seed=seed, auto_reset=False, **kw)
    rs = np.random.RandomState(seed)
    env.reset()
    for t in range(N * N // 2):
        o = env.step()
        obs0, mask0 = env.encode(0)
        live = o["done"] == 0
        eq(obs0[live], o["obs"][live], "encode(0) obs t=%d" % t)
        eq(mask0[live], o["mask"][live], "encode(0) mask t=%d" % t)
        for view in (0, 1):
            obs, mask = env.encode(view)
            empty = 2 if variant_a else 0
            eq(mask.reshape(G, N, N) != 0, obs == empty, "view %d mask == empty cells t=%d" % (view, t))
            u = rs.rand(G)
            u[:5] = [0.0, np.nextafter(1.0, 0.0), 0.5, 1.0 / 3.0, 0.999999]
            got = env.sample_actions(u, view)
            cnt = mask.sum(1).astype(np.int64)
            k = np.minimum((u * cnt).astype(np.int64), np.maximum(cnt - 1, 0))
            want = np.argsort(-mask.astype(np.int8), axis=1, kind="stable")[np.arange(G), k]
            ok = (cnt > 0) & live          # (a finished game has nobody to sample for)
            eq(got[ok], want[ok].astype(np.int32), "sample_actions view %d t=%d" % (view, t))


def raw_random_games(make, N, G, seed=0, variant_b=False):
    """BASELINE config 1: G raw HexGame instances played random-vs-random (uniformly random cells, so some moves are illegal
    and some games continue after a win, which fast_move allows), return code and the whole state compared with the oracle
    after EVERY ply (HexGame.py:85-142 / HexSingleGame.py:88-153)."""
    kind = hexref.KIND_GAME_B if variant_b else hexref.KIND_GAME_A
    env = make(kind, N, G)
    ref = hexref.RefBatch(kind, N, G)
    env.reset()
    rs = np.random.RandomState(seed)
    for t in range(N * N + 6):
        a = rs.randint(0, N * N, size=G).astype(np.int32)
        a[rs.rand(G) < 0.02] = rs.randint(-3, N * N + 3)
        eq(env.ply(a), ref.ply(a), "raw games N=%d ret t=%d" % (N, t))
        e, r = env.export(), ref.export()
        for k in ("board", "regions", "region_counter", "cur", "done", "winner"):
            eq(e[k], r[k], "raw games N=%d %s t=%d" % (N, k, t))


def golden_saturation(make, name):
    """Label-range stress (oracle/gen_golden.py: saturation_moves): both colours found ~N(N-2)/3 separate regions (region_counter
    110 / 111 on 19x19 - the packed state keeps labels in 7 bits) and then merge them all. Implementation vs the snapshots of the
    unmodified reference AND vs the oracle after every ply."""
    z = np.load(os.path.join(GOLDEN, name))
    N, moves = int(z["N"]), z["moves"]
    G = 33                                              # one chunk and a bit; every game plays the same list
    env = make(hexref.KIND_GAME_A, N, G)
    ref = hexref.RefBatch(hexref.KIND_GAME_A, N, G)
    env.reset()
    snaps = {int(t): i for i, t in enumerate(z["snap_t"])}
    top = 0
    for t, a in enumerate(moves):
        acts = np.full(G, a, np.int32)
        r = env.ply(acts)
        eq(r, ref.ply(acts), "%s ret t=%d" % (name, t))
        assert (r == z["ret"][t]).all(), (name, t)
        if t in snaps or t % 9 == 0:
            e, o = env.export(), ref.export()
            for k in ("board", "regions", "region_counter", "cur", "done", "winner"):
                eq(e[k], o[k], "%s %s t=%d" % (name, k, t))
            top = max(top, int(e["region_counter"].max()))
        if t in snaps:
            i = snaps[t]
            assert (e["regions"] == z["regions"][i].astype(np.float64)).all(), (name, "regions", t)
            assert (e["region_counter"] == z["counter"][i].astype(np.float64)).all(), (name, "counter", t)
            assert (e["board"] == z["board"][i].astype(np.float64)).all(), (name, "board", t)
    assert top >= (105 if N >= 19 else 95), top         # the label range really was exercised
    return top


def golden_preset_resets(make_raw, name, need_merges=True):
    """HexGame.__init__ with connected_stones (cached planes of HexEnv.reset, user regions=): planes adopted as they are,
    region_counter = max(plane) + 1, and the next new region takes its label from that counter. hexb_import_labels vs the
    unmodified reference's second reset."""
    z = np.load(os.path.join(GOLDEN, name))
    N, tc = int(z["N"]), z["board_true"]
    n = tc.shape[0]
    for variant, kind in (("A", hexref.KIND_GAME_A), ("B", hexref.KIND_GAME_B)):
        reg, ctr = z["regions_" + variant], z["counter_" + variant]
        env = make_raw(kind, N, n)
        env.reset()
        env.import_boards(tc)                                   # reset 0: raster-order rebuild
        e = env.export()
        eq(e["regions"], reg[:, 0].astype(np.float64), "%s %s rebuild regions" % (name, variant))
        eq(e["region_counter"], ctr[:, 0].astype(np.float64), "%s %s rebuild counter" % (name, variant))
        env.import_labels(tc, reg[:, 0])                        # reset 1: the cached planes come back
        e = env.export()
        eq(e["regions"], reg[:, 1].astype(np.float64), "%s %s adopted regions" % (name, variant))
        eq(e["region_counter"], ctr[:, 1].astype(np.float64), "%s %s adopted counter" % (name, variant))
        assert not need_merges or (ctr[:, 0] != ctr[:, 1]).any()   # the committed fixtures do contain merged presets
        mv = z["move_" + variant]
        ok = mv >= 0
        env.ply(np.where(ok, mv, 0).astype(np.int32))
        e = env.export()
        eq(e["regions"][ok], z["moved_regions_" + variant][ok].astype(np.float64), "%s %s regions after a move" % (name, variant))
        eq(e["region_counter"][ok], z["moved_counter_" + variant][ok].astype(np.float64), "%s %s counter after a move" % (name, variant))


def golden_oppredict_batched(make, name):
    """Variant-A HexEnv(opponent_policy="opponent_predict", opponent_model=..., eps=...) for a whole batch (hexb_set_opponent_eps +
    hexb_half_step) against the reference run one env per game (oppredict_*.npz: HexGame.py:165-167,354-359; the batched form of
    what scripts/selfplay.py:38-44 builds with gym.make("hex-v0", ...)). The scripted model answers for EVERY waiting game from the
    implementation's own opponent view; the implementation's draw from the game's stream decides whether random_policy moves instead.
    Pins: the draw order (rv, then random_policy's draw), the board and mask the model is shown, rewards, restarts with the opponent
    opening, the region planes."""
    from oracle.scripted import scripted_choice_a
    z = np.load(os.path.join(GOLDEN, name))
    N, seed, eps, of = int(z["N"]), int(z["seed"]), float(z["eps"]), int(z["opponent_first"])
    T, G = z["actions"].shape
    env = make(hexref.KIND_ENV_A, N, G, seed=seed, opponent_first=bool(of), manual_opponent=True)
    env.set_opponent_eps(eps)
    env.reset()

    def opponent_pass(model_calls=None, t=None):
        tm, _ = env.opp_state()
        obs1, mask1 = env.view1()
        acts = np.zeros(G, np.int32)
        for g in np.flatnonzero(tm == 1):
            acts[g] = scripted_choice_a(obs1[g], mask1[g])
            if model_calls is not None and model_calls[g]:   # the reference's model was asked in this step: same board, mask, answer
                assert np.array_equal(obs1[g], z["model_board"][t, g]), (name, t, g, "board shown to the model")
                assert np.array_equal(mask1[g], z["model_mask"][t, g]), (name, t, g, "mask shown to the model")
                assert acts[g] == z["model_action"][t, g], (name, t, g)
        return env.half_step(1, acts, want_term=True), tm == 1

    _, opened = opponent_pass()
    eq(opened, np.full(G, bool(of)), name + " who opens")
    obs, _ = env.view1()                                         # (every game is back at the agent: its view)
    eq(obs, z["obs0"], name + " obs0")
    eq(env.export()["draws"], z["draws0"], name + " draws0")
    for t in range(T):
        w = "%s t=%d " % (name, t)
        h = env.half_step(0, z["actions"][t], want_term=True)
        reward, done = h["reward"].copy(), h["done"].astype(bool)
        h, replied = opponent_pass(z["model_calls"][t], t)       # the reply to the agent's ply
        assert not (z["model_calls"][t].astype(bool) & ~replied).any(), w + "a model call without a reply"
        reward += h["reward"]
        done |= h["done"].astype(bool)
        h, _ = opponent_pass()                                   # the opening move of a game that restarted (opponent_first)
        assert not h["done"].any(), w + "an opening move cannot end a game"
        tm, _ = env.opp_state()
        assert (tm == 0).all(), w + "every game back at the agent"
        eq(reward, z["reward"][t], w + "reward")
        eq(done, z["done"][t].astype(bool), w + "done")
        obs, mask = env.view1()
        eq(obs, z["obs"][t], w + "obs")
        eq(mask, (z["obs"][t].reshape(G, -1) == 2).astype(np.uint8), w + "mask")
        e = env.export()
        eq(e["draws"], z["draws"][t], w + "draws")
        live = ~done                                             # the fixture's planes are those before the restart
        eq(e["regions"][live], z["regions"][t][live].astype(np.float64), w + "regions")
        eq(e["region_counter"][live], z["counter"][t][live].astype(np.float64), w + "counter")


def opponent_modes_fuzz(make, kind, N, G, T, seed, pool=4):
    """Split steps with random (mostly legal) moves of both sides against the oracle while the opponent modes change at random
    moments: variant B - SelfPlayEnv.set_eval switched on / off (the evaluation cycle restarts from entry 0 at every call);
    variant A - HexEnv.eps of opponent_predict re-set to a new value or switched off. Every half step: to_move, the pool entry,
    reward, done, the terminal observation; the whole exported state every few steps."""
    rs = np.random.RandomState(seed)
    C = N * N
    kw = dict(agent_mode=2) if kind == hexref.KIND_SELFPLAY_B else dict(opponent_first=bool(seed & 1))
    env = make(kind, N, G, seed=seed, manual_opponent=True, pool_size=pool, **kw)
    ref = hexref.RefBatch(kind, N, G, seed=seed, manual_opponent=True, pool_size=pool, **kw)
    env.reset(); ref.reset()
    switches = 0
    for t in range(T):
        if rs.rand() < 0.15:
            switches += 1
            if kind == hexref.KIND_SELFPLAY_B:
                flag = bool(rs.randint(2))
                env.set_eval(flag); ref.set_eval(flag)
            else:
                eps = -1.0 if rs.rand() < 0.3 else float(rs.choice([0.0, 0.25, 0.5, 1.0]))
                env.set_opponent_eps(eps); ref.set_opponent_eps(eps)
        for side in (1, 0, 1):
            tm, idx = env.opp_state()
            rtm, ridx = ref.opp_state()
            eq(tm, rtm, "to_move t=%d side=%d" % (t, side))
            eq(idx, ridx, "pool entry t=%d side=%d" % (t, side))
            obs1, mask1 = env.view1()
            robs1, rmask1 = ref.view1()
            eq(obs1[tm != 2], robs1[tm != 2], "side-to-move obs t=%d side=%d" % (t, side))
            cnt = np.maximum(mask1.sum(1), 1)
            k = (rs.rand(G) * cnt).astype(np.int64)
            acts = np.argsort(-mask1.astype(np.int8), axis=1, kind="stable")[np.arange(G), k].astype(np.int32)
            bad = rs.rand(G) < 0.01
            acts[bad] = rs.randint(-1, C + 1, size=int(bad.sum()))
            o = env.half_step(side, acts, want_term=True)
            r = ref.half_step(side, acts, want_term=True)
            for key in ("reward", "done", "to_move", "opp_index"):
                eq(o[key], r[key], "%s t=%d side=%d" % (key, t, side))
            d = r["done"].astype(bool)
            eq(o["term_obs"][d], r["term_obs"][d], "term_obs t=%d side=%d" % (t, side))
        if t % 6 == 5 or t == T - 1:
            e, re_ = env.export(), ref.export()
            for key in ("regions", "region_counter", "cur", "done", "winner", "agent", "draws"):
                eq(e[key], re_[key], "%s t=%d" % (key, t))
    eq(env.stats(), ref.stats(), "stats")
    assert switches > 0
