"""One-off CPU soak: the UNMODIFIED reference run live (oracle/gen_golden.py's drivers on /root/reference or oracle/_ref) on
many fresh seeds, every trace replayed through the C oracle AND the device logic (the emulator build of csrc/*.cuh) with the
checkers the committed fixtures go through. python tools/live_reference_soak.py [first_seed] [count] [processes]
SOAK_BIG=1: the two env rollouts on boards of 10x10 ... 19x19 (2-3 games, long enough to finish episodes)."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def one(seed):
    import numpy as np
    import parity
    from oracle import gen_golden as gg, hexref
    from test_emu_parity import make as emu, EmuBatch
    from test_oracle_golden import _check_rollout
    import test_oracle_golden as tog
    rs = np.random.RandomState(seed)
    oracle = lambda kind, N, G, **kw: hexref.RefBatch(kind, N, G, **kw)
    done = []
    stock_load = tog.load
    with tempfile.TemporaryDirectory() as tmp:
        gg.OUT = tmp
        try:
            big = os.environ.get("SOAK_BIG") == "1"     # large boards: few games, rollouts long enough to finish episodes
            N = int(rs.randint(12, 20)) if big else int(rs.randint(3, 12))
            G = int(rs.randint(2, 4)) if big else int(rs.randint(2, 9))
            T = int(N * N // 2 + rs.randint(10, 60)) if big else int(rs.randint(20, 5 * N + 30))
            am, fused = int(rs.randint(3)), bool(rs.randint(2))
            o = gg.rollout("B", N, G, T, seed=seed, agent_mode=am, fused=fused)
            p = os.path.join(tmp, "selfplay_live.npz")
            np.savez_compressed(p, N=N, seed=seed, agent_mode=am, fused=int(fused), **o)
            _check_rollout(p, hexref.KIND_SELFPLAY_B); parity.golden_rollout(emu, p)
            done.append("selfplay N=%d" % N)
            N = int(rs.randint(10, 20)) if big else int(rs.randint(3, 10))
            of = int(rs.randint(2))
            o = gg.rollout("A", N, G, T, seed=seed + 1, agent_mode=0, fused=fused, opponent_first=bool(of))
            p = os.path.join(tmp, "envA_live.npz")
            np.savez_compressed(p, N=N, seed=seed + 1, opponent_first=of, fused=int(fused), **o)
            _check_rollout(p, hexref.KIND_ENV_A); parity.golden_rollout(emu, p)
            done.append("envA N=%d" % N)
            N = int(rs.randint(3, 8)); pool = int(rs.randint(1, 6))
            sched = {int(rs.randint(2, T // 2)): True, int(rs.randint(T // 2, T)): False} if rs.randint(2) else None
            o = gg.rollout_scripted_opponent(N, G, T, seed=seed + 2, agent_mode=am, pool=pool, eval_schedule=sched)
            p = os.path.join(tmp, "oppmodel_live.npz")
            np.savez_compressed(p, N=N, seed=seed + 2, agent_mode=am, pool=pool, **o)
            for mk in (oracle, emu):
                parity.golden_oppmodel(mk, p)
            done.append("oppmodel N=%d pool=%d" % (N, pool))
            N = int(rs.randint(3, 7)); eps = float(rs.choice([0.0, 0.2, 0.5, 0.8, 1.0]))
            gg.gen_opponent_predict(N, G, min(T, 40), seed=seed + 3, eps=eps, opponent_first=bool(of))
            for mk in (oracle, emu):
                parity.golden_oppredict_batched(mk, os.path.join(tmp, "oppredict_N%d_of%d.npz" % (N, of)))
            done.append("oppredict N=%d eps=%.1f" % (N, eps))
            for variant in ("A", "B"):
                N = int(rs.randint(3, 14))
                gg.gen_raw_games(variant, N, 2, seed + 4)
                name = "game_%s_N%d.npz" % (variant, N)
                p = os.path.join(tmp, name)
                tog.load = lambda _n, p=p: np.load(p)
                tog.test_raw_game_traces(name)
                if variant == "A":
                    parity.golden_raw_game(emu, p)
                done.append("raw %s N=%d" % (variant, N))
            # preset boards: HexGame.__init__ rebuilding the planes in raster order, and HexEnv.reset adopting its cached planes
            N = int(rs.randint(3, 14)); n = int(rs.randint(2, 7))
            gg.gen_preset_boards(N, n, seed + 5)
            p = os.path.join(tmp, "preset_N%d.npz" % N)
            tc = np.load(p)["board_true"]

            def oracle_raw(kind, N, G):
                b = hexref.RefBatch(kind, N, G)
                b.set_board(tc if kind == hexref.KIND_GAME_A else np.where(tc == 0, -1, np.where(tc == 1, 1, 0)).astype(np.int8), cur=0)
                return b

            def emu_raw(kind, N, G):
                env = EmuBatch(0 if kind == hexref.KIND_GAME_A else 1, N, G, raw=True)
                env.reset()
                env.import_boards(tc, np.zeros(G, np.int8))
                return env
            parity.golden_preset(oracle_raw, p); parity.golden_preset(emu_raw, p)
            gg.gen_preset_resets(N, n, seed + 6)
            parity.golden_preset_resets(lambda kind, N, G: EmuBatch(0 if kind == hexref.KIND_GAME_A else 1, N, G, raw=True),
                                        os.path.join(tmp, "presetreset_N%d.npz" % N), need_merges=False)
            done.append("presets N=%d" % N)
        except Exception as e:   # noqa: BLE001 - report the seed and go on
            return seed, "FAIL after %s: %s: %s" % (done, type(e).__name__, str(e)[:300])
        finally:
            tog.load = stock_load
    return seed, None


if __name__ == "__main__":
    from multiprocessing import Pool
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    procs = int(sys.argv[3]) if len(sys.argv) > 3 else (os.cpu_count() or 1)
    bad = 0
    with Pool(procs) as pool:
        for seed, err in pool.imap_unordered(one, range(first, first + 10 * count, 10)):
            if err:
                bad += 1
                print("seed %d: %s" % (seed, err), flush=True)
    print("live-reference soak done: %d seeds x 7 traces (env rollouts of both variants, learned opponents + evaluation cycle, "
          "batched opponent_predict, raw games of both variants, preset boards and repeated resets on them), failures: %d" % (count, bad))
