"""TEST INFRASTRUCTURE ONLY - generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden
The reference is imported read-only through oracle/ref_harness.py (stubs for gymnasium/pygame,
per-game Philox stream in place of the global ``random``). The .npz files are the pins for the
C oracle (tests/test_oracle_golden.py) and, through the oracle and directly, for the CUDA path
(tests/test_gpu_parity.py). Nothing here runs on the GPU box.

Fixture families
  game_{A,B}_N*.npz      raw HexGame.make_move traces (random moves incl. occupied cells, played on
                         past the win until the board is full): ret/board/regions/counter/cur/done/winner per ply
  selfplay_N*_a*.npz     SelfPlayEnv (variant B) + BaseRandomPolicy rollouts with DummyVecEnv-style auto-reset;
                         agent either sampled by BaseRandomPolicy from the same stream ("fused") or given
                         externally (with some illegal moves)
  envA_N*_of*.npz        variant-A HexEnv(opponent_policy=minihex.random_policy) rollouts, same two agent modes
  oppmodel_N*_a*.npz     SelfPlayEnv with OpponentPolicy opponents (scripted stand-ins for SB3 models, oracle/scripted.py): the
                         learned-opponent path incl. the 80/20 best/pool choice of setup_opponents and the opponent's observation
  evalpool_N*_a*.npz     the same with SelfPlayEnv.set_eval switched on and off in mid-run: the evaluation cycle through the pool
                         (setup_opponents in eval_state: episode k meets opponent_models[k], nothing is drawn)
  preset_N*.npz          HexGame.__init__ on preset boards (raster-order region rebuild), both variants
  kat.npz                the four hand-checked known-answer tests of SURVEY.md section 8c
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as rh  # noqa: E402
from oracle.philox import GameStream, ListStream  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
NONE = -1


def code(w):
    return NONE if w is None else int(w)


def snap_game(g):
    return (np.array(g.board, dtype=np.int8).copy(), np.array(g.regions, dtype=np.uint8).copy(),
            np.array(g.region_counter, dtype=np.int16).copy(), int(g.current_player_num), int(bool(g.done)), code(g.winner))


def gen_raw_games(variant, N, n_games, seed):
    minihex, A, B, S = rh.load()
    rs = np.random.RandomState(seed)
    C = N * N
    T = C + C // 2 + 4
    moves = np.zeros((n_games, T), np.int32)
    ret = np.zeros((n_games, T), np.int8)
    board = np.zeros((n_games, T, N, N), np.int8)
    regions = np.zeros((n_games, T, 2, N + 2, N + 2), np.uint8)
    counter = np.zeros((n_games, T, 2), np.int16)
    cur = np.zeros((n_games, T), np.int8)
    done = np.zeros((n_games, T), np.uint8)
    winner = np.zeros((n_games, T), np.int8)
    for gi in range(n_games):
        if variant == "A":
            g = A.HexGame(A.player.BLACK, A.player.EMPTY * np.ones((N, N)), A.player.BLACK)
        else:
            g = B.HexGame(0, np.zeros((N, N)))
        # a random permutation with ~1/3 random repeats mixed in (=> some moves hit occupied cells)
        perm = list(rs.permutation(C))
        seq = []
        while len(seq) < T:
            if perm and rs.rand() > 0.3:
                seq.append(perm.pop())
            else:
                seq.append(int(rs.randint(C)))
        for t, a in enumerate(seq):
            moves[gi, t] = a
            ret[gi, t] = code(g.make_move(a))
            b_, r_, c_, cu, dn, wn = snap_game(g)
            board[gi, t], regions[gi, t], counter[gi, t], cur[gi, t], done[gi, t], winner[gi, t] = b_, r_, c_, cu, dn, wn
    np.savez_compressed(os.path.join(OUT, "game_%s_N%d.npz" % (variant, N)), N=N, moves=moves, ret=ret, board=board,
                        regions=regions, counter=counter, cur=cur, done=done, winner=winner)


def rollout(kind, N, G, T, seed, agent_mode, fused, opponent_first=False, illegal_rate=0.05):
    """kind 'B' = SelfPlayEnv, 'A' = variant-A HexEnv. One env object per game, own stream each."""
    minihex, A, B, S = rh.load()
    rs = np.random.RandomState(seed ^ 0x5EED)
    C = N * N
    out = dict(actions=np.zeros((T, G), np.int32), obs=np.zeros((T, G, N, N), np.int8), mask=np.zeros((T, G, C), np.uint8),
               reward=np.zeros((T, G), np.float32), done=np.zeros((T, G), np.uint8), term_obs=np.zeros((T, G, N, N), np.int8),
               regions=np.zeros((T, G, 2, N + 2, N + 2), np.uint8), counter=np.zeros((T, G, 2), np.int16),
               sim_cur=np.zeros((T, G), np.int8), draws=np.zeros((T, G), np.uint32),
               obs0=np.zeros((G, N, N), np.int8), mask0=np.zeros((G, C), np.uint8), agent=np.zeros(G, np.int8),
               draws0=np.zeros(G, np.uint32))
    for gi in range(G):
        stream = GameStream(seed, gi)
        rh.set_rng(stream)
        if kind == "B":
            env = S.selfplay_wrapper(B.HexEnv)(board_size=N, agent_player_num=None if agent_mode == 2 else agent_mode)
            mask_fn = env.legal_actions
            choose = lambda board: S.BaseRandomPolicy().choose_action(board)
        else:
            env = A.HexEnv(opponent_policy=minihex.random_policy, board_size=N,
                           current_player_num=A.player.WHITE if opponent_first else A.player.BLACK)
            mask_fn = env.get_action_mask
            choose = lambda board: minihex.random_policy(board)
        obs, _ = env.reset()
        out["obs0"][gi] = obs
        out["mask0"][gi] = mask_fn()
        out["agent"][gi] = env.agent_player_num if kind == "B" else 0
        out["draws0"][gi] = stream.idx
        for t in range(T):
            mask = mask_fn()
            if fused:
                a = int(choose(obs))
            else:
                legal = np.flatnonzero(mask)
                a = int(rs.randint(C)) if rs.rand() < illegal_rate else int(legal[rs.randint(len(legal))])
            obs, r, done, _, _ = env.step(a)
            out["actions"][t, gi] = a
            out["reward"][t, gi] = r
            out["done"][t, gi] = done
            if done:
                out["term_obs"][t, gi] = obs
                obs, _ = env.reset()
            out["obs"][t, gi] = obs
            out["mask"][t, gi] = mask_fn()
            out["regions"][t, gi] = env.simulator.regions
            out["counter"][t, gi] = env.simulator.region_counter
            out["sim_cur"][t, gi] = env.simulator.current_player_num
            out["draws"][t, gi] = stream.idx
    return out


def rollout_scripted_opponent(N, G, T, seed, agent_mode, pool, illegal_rate=0.04, eval_schedule=None):
    """SelfPlayEnv whose opponents are OpponentPolicy objects (SelfplayWrapper.py:26-35) around scripted models: the
    learned-opponent path (setup_opponents' 80/20 choice, continue_game with action masks). One env per game.
    eval_schedule {t: flag}: env.set_eval(flag) right before step t (SelfplayWrapper.py:117-120, what SelfPlayCallback does around
    its evaluation, EvaluationCallback.py:31-33): while it is on, setup_opponents hands episode k since the call
    opponent_models[k] (:92-96) and draws nothing; recorded as eval_at[t] (-1 = no call)."""
    from oracle.scripted import ScriptedModel
    minihex, A, B, S = rh.load()
    rs = np.random.RandomState(seed ^ 0xBEEF)
    C = N * N
    out = dict(actions=np.zeros((T, G), np.int32), opp_actions=-np.ones((T, G, 2), np.int32), opp_model=-9 * np.ones((T, G, 2), np.int32),
               opp0_action=-np.ones(G, np.int32), opp0_model=-9 * np.ones(G, np.int32),
               obs=np.zeros((T, G, N, N), np.int8), mask=np.zeros((T, G, C), np.uint8), reward=np.zeros((T, G), np.float32),
               done=np.zeros((T, G), np.uint8), term_obs=np.zeros((T, G, N, N), np.int8),
               regions=np.zeros((T, G, 2, N + 2, N + 2), np.uint8), counter=np.zeros((T, G, 2), np.int16),
               sim_cur=np.zeros((T, G), np.int8), draws=np.zeros((T, G), np.uint32), obs0=np.zeros((G, N, N), np.int8),
               mask0=np.zeros((G, C), np.uint8), agent=np.zeros(G, np.int8), draws0=np.zeros(G, np.uint32))
    if eval_schedule:
        out["eval_at"] = -np.ones(T, np.int8)
        for t, flag in eval_schedule.items():
            out["eval_at"][t] = int(bool(flag))
    for gi in range(G):
        stream = GameStream(seed, gi)
        rh.set_rng(stream)
        log = []
        env = S.selfplay_wrapper(B.HexEnv)(base_model=ScriptedModel(-1, log), scores=np.zeros(pool), board_size=N, buffer_size=pool,
                                           agent_player_num=None if agent_mode == 2 else agent_mode)
        for k in range(pool):
            env.opponent_models[k] = S.OpponentPolicy(ScriptedModel(k, log))
        obs, _ = env.reset()
        if log:
            out["opp0_model"][gi], out["opp0_action"][gi] = log[0]
        del log[:]
        out["obs0"][gi], out["mask0"][gi], out["agent"][gi], out["draws0"][gi] = obs, env.legal_actions(), env.agent_player_num, stream.idx
        for t in range(T):
            if eval_schedule and t in eval_schedule:
                env.set_eval(bool(eval_schedule[t]))
            legal = np.flatnonzero(env.legal_actions())
            a = int(rs.randint(C)) if rs.rand() < illegal_rate else int(legal[rs.randint(len(legal))])
            obs, r, done, _, _ = env.step(a)
            out["actions"][t, gi], out["reward"][t, gi], out["done"][t, gi] = a, r, done
            if done:
                out["term_obs"][t, gi] = obs
                obs, _ = env.reset()
            assert len(log) <= 2
            for j, (ident, oa) in enumerate(log):
                out["opp_model"][t, gi, j], out["opp_actions"][t, gi, j] = ident, oa
            del log[:]
            out["obs"][t, gi], out["mask"][t, gi] = obs, env.legal_actions()
            out["regions"][t, gi], out["counter"][t, gi] = env.simulator.regions, env.simulator.region_counter
            out["sim_cur"][t, gi], out["draws"][t, gi] = env.simulator.current_player_num, stream.idx
    return out


def gen_preset_boards(N, n, seed):
    """HexGame.__init__ with a preset board and connected_stones=None (HexGame.py:53-61, HexSingleGame.py:57-65): the planes
    are rebuilt by flood_fill in raster order. Random boards (true coordinates), both variants."""
    minihex, A, B, S = rh.load()
    rs = np.random.RandomState(seed)
    true_codes = rs.choice([0, 1, 2], size=(n, N, N), p=[0.3, 0.3, 0.4]).astype(np.int8)
    out = dict(board_true=true_codes, regions_A=np.zeros((n, 2, N + 2, N + 2), np.uint8), counter_A=np.zeros((n, 2), np.int16),
               regions_B=np.zeros((n, 2, N + 2, N + 2), np.uint8), counter_B=np.zeros((n, 2), np.int16))
    for i in range(n):
        ga = A.HexGame(A.player.BLACK, true_codes[i].astype(np.float64), A.player.BLACK)
        bb = np.where(true_codes[i] == 0, -1.0, np.where(true_codes[i] == 1, 1.0, 0.0))
        gb = B.HexGame(0, bb)
        out["regions_A"][i], out["counter_A"][i] = ga.regions, ga.region_counter
        out["regions_B"][i], out["counter_B"][i] = gb.regions, gb.region_counter
    np.savez_compressed(os.path.join(OUT, "preset_N%d.npz" % N), N=N, **out)


def gen_kats():
    """SURVEY.md section 8c KAT-1..4, re-derived from the reference here and stored verbatim."""
    minihex, A, B, S = rh.load()
    k = {}
    g = A.HexGame(A.player.BLACK, A.player.EMPTY * np.ones((3, 3)), A.player.BLACK)
    rets, empties = [], []
    for m in [4, 0, 1, 3, 7]:
        rets.append(code(g.make_move(m)))
        empties.append(int(g.empty_fields))
    k["kat1_ret"], k["kat1_empty_fields"] = np.array(rets), np.array(empties)
    k["kat1_board"], k["kat1_regions"], k["kat1_counter"] = g.board.copy(), g.regions.copy(), g.region_counter.copy()
    k["kat1_again"] = np.array([code(g.make_move(4)), g.current_player_num])

    def selfplay(draws, agent, acts, name):
        rh.set_rng(ListStream(draws))
        env = S.selfplay_wrapper(B.HexEnv)(board_size=3, agent_player_num=agent)
        obs, _ = env.reset()
        seq_obs, seq_r, seq_d = [np.array(obs)], [], []
        for a in acts:
            obs, r, d, _, _ = env.step(a)
            seq_obs.append(np.array(obs)); seq_r.append(r); seq_d.append(d)
        k[name + "_obs"], k[name + "_r"], k[name + "_d"] = np.array(seq_obs), np.array(seq_r), np.array(seq_d)
        k[name + "_regions"] = env.simulator.regions.copy()
        k[name + "_envcur_simcur_winner"] = np.array([env.current_player_num, env.simulator.current_player_num, code(env.winner)])

    selfplay([0.1, 0.2, 0.0, 0.3, 0.99], 0, [4, 1, 7], "kat2")
    selfplay([0.1, 0.2, 0.5, 0.5, 0.5], 1, [0, 0], "kat3")
    rh.set_rng(ListStream([0.5, 0.5, 0.5]))
    env = A.HexEnv(opponent_policy=minihex.random_policy, board_size=3)
    env.reset()
    o1, r1, d1, _, i1 = env.step(4)
    o1 = np.array(o1)
    o2, r2, d2, _, i2 = env.step(4)
    k["kat4_obs"] = np.array([o1, np.array(o2)])
    k["kat4_r"], k["kat4_d"] = np.array([r1, r2]), np.array([d1, d2])
    k["kat4_last_move_opponent"] = np.array([i1["last_move_opponent"]])
    k["kat4_winner"] = np.array([code(i2["winner"])])
    np.savez_compressed(os.path.join(OUT, "kat.npz"), **k)


def saturation_moves(N):
    """Label-range stress: both colours found as many separate regions as the board allows before anything merges.
    Cells with (x - y) % 3 == k form an independent set of the hex adjacency (a 3-colouring class). BLACK takes class 0 on its
    interior rows 1..N-2, WHITE class 1 on its interior columns 1..N-2: every such stone is isolated, so region_counter reaches
    3 + ~N(N-2)/3 for both colours (111 on 19x19; labels must stay below 128 in the packed state). Then the remaining cells are
    filled in raster order, every stone merging up to three regions, until the board is full (play continues past the win).
    Moves alternate BLACK / WHITE and never hit an occupied cell."""
    blacks = [y * N + x for y in range(1, N - 1) for x in range(N) if (x - y) % 3 == 0]
    whites = [y * N + x for y in range(N) for x in range(1, N - 1) if (x - y) % 3 == 1]
    taken = set(blacks) | set(whites)
    rest = [c for c in range(N * N) if c not in taken]
    moves, bi, wi, ri = [], 0, 0, 0
    for ply in range(N * N):
        if ply % 2 == 0 and bi < len(blacks):
            moves.append(blacks[bi]); bi += 1
        elif ply % 2 == 1 and wi < len(whites):
            moves.append(whites[wi]); wi += 1
        elif ri < len(rest):
            moves.append(rest[ri]); ri += 1
        elif bi < len(blacks):
            moves.append(blacks[bi]); bi += 1
        else:
            moves.append(whites[wi]); wi += 1
    assert sorted(moves) == list(range(N * N))
    return moves


def gen_saturation(N):
    """Variant-A HexGame driven through saturation_moves: return codes and snapshots of the planes / counters from the reference."""
    minihex, A, B, S = rh.load()
    moves = saturation_moves(N)
    g = A.HexGame(A.player.BLACK, A.player.EMPTY * np.ones((N, N)), A.player.BLACK)
    ret, snaps_t, regions, counter, board = [], [], [], [], []
    for t, a in enumerate(moves):
        ret.append(code(g.make_move(a)))
        if t % 16 == 15 or t >= len(moves) - 3 or t == 2 * (N * (N - 2) // 3):
            snaps_t.append(t)
            regions.append(np.array(g.regions, dtype=np.uint8))
            counter.append(np.array(g.region_counter, dtype=np.int16))
            board.append(np.array(g.board, dtype=np.int8))
    counter = np.array(counter)
    assert counter.max() >= (105 if N >= 19 else 3 + (N - 2) * N // 3 - 2), counter.max()
    np.savez_compressed(os.path.join(OUT, "saturation_N%d.npz" % N), N=N, moves=np.array(moves, np.int32), ret=np.array(ret, np.int8),
                        snap_t=np.array(snaps_t, np.int32), regions=np.array(regions), counter=counter, board=np.array(board))


def gen_facade(N, seed):
    """The small HexGame methods (a2 / a3 / a7): is_valid_move, action_to_coordinate, coordinate_to_action, get_possible_actions
    at positions along a random game, both variants; which out-of-range actions raise IndexError."""
    minihex, A, B, S = rh.load()
    rs = np.random.RandomState(seed)
    C = N * N
    out = {}
    for variant in ("A", "B"):
        if variant == "A":
            g = A.HexGame(A.player.BLACK, A.player.EMPTY * np.ones((N, N)), A.player.BLACK)
        else:
            g = B.HexGame(0, np.zeros((N, N)))
        moves = np.zeros(C - 2, np.int64)
        valid, possible, npossible, at = [], [], [], []
        for t in range(C - 2):
            # a legal cell of the CURRENT view (after an illegal move HexEnv ends the episode; a raw game that plays on has its
            # board and planes in different coordinate systems in the reference, which is outside every caller's contract)
            legal = [k for k in range(C) if g.is_valid_move(k)]
            a = moves[t] = legal[rs.randint(len(legal))]
            if t % 3 == 0:
                at.append(t)
                valid.append([bool(g.is_valid_move(k)) for k in range(C)])
                pa = np.asarray(g.get_possible_actions())
                npossible.append(len(pa))
                possible.append(np.concatenate([pa, -np.ones(C - len(pa), pa.dtype)]))
            if variant == "B":
                # raw variant-B games are driven the way HexEnv drives them: the board is flipped after every ply, the
                # action is a cell of the mover's view (oracle/gen_golden.py: gen_raw_games)
                pass
            g.make_move(int(a))
            if variant == "B":
                bb = g.board.copy().T
                bb[bb == -1] = 2; bb[bb == 1] = -1; bb[bb == 2] = 1
                g.board = bb
        out["moves_" + variant] = moves.astype(np.int32)
        out["at_" + variant] = np.array(at, np.int32)
        out["valid_" + variant] = np.array(valid, np.uint8)
        out["possible_" + variant] = np.array(possible, np.int32)
        out["npossible_" + variant] = np.array(npossible, np.int32)
        raises = []
        for k in (C, C + 1, C + N, 2 * C):
            try:
                g.is_valid_move(k)
                raises.append(0)
            except IndexError:
                raises.append(1)
        out["oob_actions"] = np.array([C, C + 1, C + N, 2 * C], np.int32)
        out["oob_raises_" + variant] = np.array(raises, np.uint8)
        out["coords_" + variant] = np.array([g.action_to_coordinate(k) for k in range(C)], np.int32)
        out["actions_of_coords_" + variant] = np.array([g.coordinate_to_action(tuple(g.action_to_coordinate(k))) for k in range(C)], np.int32)
    np.savez_compressed(os.path.join(OUT, "facade_N%d.npz" % N), N=N, **out)


def gen_opponent_predict(N, G, T, seed, eps, opponent_first):
    """Variant-A HexEnv with opponent_policy="opponent_predict" (HexGame.py:165-167, 354-359): with probability eps the opponent
    is random_policy, else the model's deterministic prediction on the inverted board with the mask of that view."""
    from oracle.scripted import ScriptedModelA
    minihex, A, B, S = rh.load()
    rs = np.random.RandomState(seed ^ 0xACE)
    C = N * N
    out = dict(actions=np.zeros((T, G), np.int32), obs=np.zeros((T, G, N, N), np.int8), reward=np.zeros((T, G), np.float32),
               done=np.zeros((T, G), np.uint8), last_move_opponent=-np.ones((T, G), np.int32), winner=-9 * np.ones((T, G), np.int8),
               draws=np.zeros((T, G), np.uint32), model_calls=np.zeros((T, G), np.int32), model_action=-np.ones((T, G), np.int32),
               model_mask=np.zeros((T, G, C), np.uint8), model_board=np.zeros((T, G, N, N), np.int8),
               obs0=np.zeros((G, N, N), np.int8), draws0=np.zeros(G, np.uint32), regions=np.zeros((T, G, 2, N + 2, N + 2), np.uint8),
               counter=np.zeros((T, G, 2), np.int16))
    for gi in range(G):
        stream = GameStream(seed, gi)
        rh.set_rng(stream)
        log = []
        env = A.HexEnv(opponent_policy="opponent_predict", opponent_model=ScriptedModelA(log), board_size=N, eps=eps,
                       current_player_num=A.player.WHITE if opponent_first else A.player.BLACK)
        obs, _ = env.reset()
        del log[:]
        out["obs0"][gi], out["draws0"][gi] = obs, stream.idx
        for t in range(T):
            legal = np.flatnonzero(env.get_action_mask())
            a = int(rs.randint(C)) if rs.rand() < 0.04 else int(legal[rs.randint(len(legal))])
            obs, r, done, _, info = env.step(a)
            out["actions"][t, gi], out["reward"][t, gi], out["done"][t, gi] = a, r, done
            out["last_move_opponent"][t, gi] = -1 if info["last_move_opponent"] is None else int(info["last_move_opponent"])
            out["winner"][t, gi] = code(info["winner"])
            out["model_calls"][t, gi] = len(log)
            if log:
                out["model_action"][t, gi], out["model_mask"][t, gi], out["model_board"][t, gi] = log[-1]
            del log[:]
            out["regions"][t, gi], out["counter"][t, gi] = env.simulator.regions, env.simulator.region_counter
            if done:
                obs, _ = env.reset()
                del log[:]
            out["obs"][t, gi], out["draws"][t, gi] = obs, stream.idx
    np.savez_compressed(os.path.join(OUT, "oppredict_N%d_of%d.npz" % (N, int(opponent_first))), N=N, seed=seed, eps=eps,
                        opponent_first=int(opponent_first), **out)


def gen_preset_resets(N, n, seed):
    """HexEnv.reset called repeatedly on a preset board (both variants): the first reset rebuilds the planes in raster order, the
    later ones adopt the cached planes with region_counter = max(plane) + 1 (HexGame.py:207-220 / HexSingleGame.py:211-231), which
    differs from the first reset's counters when the rebuild merged regions. Also user-supplied regions= at construction."""
    minihex, A, B, S = rh.load()
    rs = np.random.RandomState(seed)
    true_codes = rs.choice([0, 1, 2], size=(n, N, N), p=[0.3, 0.3, 0.4]).astype(np.int8)
    out = dict(board_true=true_codes)
    for variant in ("A", "B"):
        reg = np.zeros((n, 3, 2, N + 2, N + 2), np.uint8)
        ctr = np.zeros((n, 3, 2), np.int16)
        moved = np.zeros((n, 2, N + 2, N + 2), np.uint8)
        moved_ctr = np.zeros((n, 2), np.int16)
        move = -np.ones(n, np.int32)
        for i in range(n):
            if variant == "A":
                env = A.HexEnv(opponent_policy=None, board=true_codes[i].astype(np.float64), board_size=N)
                make = lambda regions: A.HexEnv(opponent_policy=None, board=true_codes[i].astype(np.float64), regions=regions, board_size=N)
                empty = 2
            else:
                bb = np.where(true_codes[i] == 0, -1.0, np.where(true_codes[i] == 1, 1.0, 0.0))
                env = B.HexEnv(board=bb, board_size=N)
                make = lambda regions: B.HexEnv(board=bb, regions=regions, board_size=N)
                empty = 0
            for k in range(2):       # reset 0 rebuilds, reset 1 adopts the cached planes
                env.reset()
                reg[i, k], ctr[i, k] = env.simulator.regions, env.simulator.region_counter
            env2 = make(np.array(reg[i, 0], dtype=np.float64))   # user-supplied regions= : adopted from the first reset on
            env2.reset()
            reg[i, 2], ctr[i, 2] = env2.simulator.regions, env2.simulator.region_counter
            free = np.flatnonzero(env.simulator.board.flatten() == empty)
            if len(free):            # one move on the adopted state: new regions take their label from the adopted counter
                move[i] = int(free[rs.randint(len(free))])
                env.simulator.make_move(int(move[i]))
                moved[i], moved_ctr[i] = env.simulator.regions, env.simulator.region_counter
        out["regions_" + variant], out["counter_" + variant] = reg, ctr
        out["move_" + variant], out["moved_regions_" + variant], out["moved_counter_" + variant] = move, moved, moved_ctr
    np.savez_compressed(os.path.join(OUT, "presetreset_N%d.npz" % N), N=N, **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    gen_kats()
    for N in (18, 19):
        gen_saturation(N)
    for N in (4, 7):
        gen_facade(N, 500 + N)
    for N, G, T in [(4, 12, 30), (7, 8, 60)]:
        for of in (False, True):
            gen_opponent_predict(N, G, T, seed=4000 + N, eps=0.5, opponent_first=of)
    for N, n in [(4, 40), (7, 20)]:
        gen_preset_resets(N, n, 600 + N)
    for N, n in [(3, 12), (4, 8), (5, 8), (7, 4), (11, 3), (13, 1)]:
        gen_raw_games("A", N, n, 100 + N)
        gen_raw_games("B", N, n, 200 + N)
    for N, G, T in [(3, 24, 24), (5, 16, 48), (6, 12, 48), (11, 8, 140)]:
        for agent_mode in (0, 1, 2):
            for fused in (1, 0):
                o = rollout("B", N, G, T, seed=1000 + N, agent_mode=agent_mode, fused=bool(fused))
                np.savez_compressed(os.path.join(OUT, "selfplay_N%d_a%d_f%d.npz" % (N, agent_mode, fused)), N=N, seed=1000 + N,
                                    agent_mode=agent_mode, fused=fused, **o)
    for N, G, T in [(3, 24, 24), (5, 16, 40), (7, 12, 60)]:
        for of in (0, 1):
            for fused in (1, 0):
                o = rollout("A", N, G, T, seed=2000 + N, agent_mode=0, fused=bool(fused), opponent_first=bool(of))
                np.savez_compressed(os.path.join(OUT, "envA_N%d_of%d_f%d.npz" % (N, of, fused)), N=N, seed=2000 + N,
                                    opponent_first=of, fused=fused, **o)
    for N, n in [(4, 30), (7, 20), (11, 12)]:
        gen_preset_boards(N, n, 400 + N)
    for N, G, T, pool in [(4, 16, 30, 3), (7, 10, 70, 5), (11, 6, 120, 20)]:
        for agent_mode in (0, 1, 2):
            o = rollout_scripted_opponent(N, G, T, seed=3000 + N, agent_mode=agent_mode, pool=pool)
            np.savez_compressed(os.path.join(OUT, "oppmodel_N%d_a%d.npz" % (N, agent_mode)), N=N, seed=3000 + N, agent_mode=agent_mode,
                                pool=pool, **o)
    # evaluation mode switched on and off in mid-run; long enough for every game to outlast the pool (the index then stays put)
    for N, G, T, pool, sched in [(4, 16, 90, 3, {20: True, 65: False}), (5, 12, 150, 6, {0: True, 100: False, 120: True})]:
        for agent_mode in (0, 2):
            o = rollout_scripted_opponent(N, G, T, seed=3500 + N, agent_mode=agent_mode, pool=pool, eval_schedule=sched)
            np.savez_compressed(os.path.join(OUT, "evalpool_N%d_a%d.npz" % (N, agent_mode)), N=N, seed=3500 + N, agent_mode=agent_mode,
                                pool=pool, **o)
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("golden fixtures: %d files, %.1f KiB" % (len(os.listdir(OUT)), total / 1024.0))


if __name__ == "__main__":
    main()
