"""TEST INFRASTRUCTURE ONLY - ctypes wrapper around oracle/hexref.c (the CPU restatement of the reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KIND_ENV_A = 0       # minihex/HexGame.py::HexEnv with opponent_policy=minihex.random_policy
KIND_SELFPLAY_B = 1  # selfplay_wrapper(minihex/HexSingleGame.py::HexEnv) with BaseRandomPolicy
KIND_GAME_A = 2      # raw minihex/HexGame.py::HexGame
KIND_GAME_B = 3      # raw minihex/HexSingleGame.py::HexGame


def build(force=False):
    so = os.path.join(_HERE, "libhexref.so")
    src = os.path.join(_HERE, "hexref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libhexref.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        vp, i32, i64, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64
        L.hexref_batch_create.restype = vp
        L.hexref_batch_create.argtypes = [i32, i32, i64, u64, i64, i32, i32, i32]
        L.hexref_batch_destroy.argtypes = [vp]
        L.hexref_batch_reset.argtypes = [vp] + [vp] * 4
        L.hexref_batch_step.argtypes = [vp, vp, vp, i32] + [vp] * 6
        L.hexref_batch_ply.argtypes = [vp, vp, vp]
        L.hexref_batch_export.argtypes = [vp] + [vp] * 8
        L.hexref_batch_stats.argtypes = [vp, vp]
        L.hexref_batch_set_board.argtypes = [vp, vp, i32]
        L.hexref_draw.restype = ctypes.c_double
        L.hexref_draw.argtypes = [u64, u64, ctypes.c_uint32]
        L.hexref_philox4x32_10.argtypes = [vp, vp, vp]
        L.hexref_set_threads.argtypes = [i32]
        L.hexref_batch_set_manual.argtypes = [vp, i32]
        L.hexref_batch_set_eval.argtypes = [vp, i32]
        L.hexref_batch_set_opponent_eps.argtypes = [vp, ctypes.c_double]
        L.hexref_batch_half_step.argtypes = [vp, i32, vp, i32] + [vp] * 5
        L.hexref_batch_observe.argtypes = [vp, vp, vp]
        L.hexref_batch_observe_opponent.argtypes = [vp, vp, vp]
        L.hexref_batch_opp_state.argtypes = [vp, vp, vp]
        L.hexref_batch_info.argtypes = [vp, vp, vp]
        L.hexref_batch_env_set_board.argtypes = [vp, vp, vp]
        L.hexref_batch_set_board_labels.argtypes = [vp, vp, vp, i32]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class RefBatch(object):
    """G independent reference environments (or raw games) stepped in a loop on the CPU."""

    def __init__(self, kind, board_size, num_games, seed=0, game_offset=0, agent_mode=0,
                 opponent_first=False, eval_state=False, manual_opponent=False, pool_size=0):
        self.kind, self.N, self.G = kind, board_size, num_games
        self.C = board_size * board_size
        self._h = lib().hexref_batch_create(kind, board_size, num_games, seed, game_offset, agent_mode,
                                            int(opponent_first), int(eval_state))
        if not self._h:
            raise ValueError("bad oracle config")
        if manual_opponent:
            self.set_manual_opponent(pool_size)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().hexref_batch_destroy(self._h)
            self._h = None

    def reset(self, reset_mask=None, open_u=None):
        obs = np.empty((self.G, self.N, self.N), np.int8)
        mask = np.empty((self.G, self.C), np.uint8)
        rm = None if reset_mask is None else np.ascontiguousarray(reset_mask, np.uint8)
        ou = None if open_u is None else np.ascontiguousarray(open_u, np.float64)
        lib().hexref_batch_reset(self._h, _p(rm), _p(ou), _p(obs), _p(mask))
        return obs, mask

    def step(self, actions=None, opp_u=None, auto_reset=True, want_term=False):
        obs = np.empty((self.G, self.N, self.N), np.int8)
        mask = np.empty((self.G, self.C), np.uint8)
        reward = np.empty(self.G, np.float32)
        done = np.empty(self.G, np.uint8)
        term = np.zeros((self.G, self.N, self.N), np.int8) if want_term else None
        aout = np.empty(self.G, np.int32)
        a = None if actions is None else np.ascontiguousarray(actions, np.int32)
        u = None if opp_u is None else np.ascontiguousarray(opp_u, np.float64)
        lib().hexref_batch_step(self._h, _p(a), _p(u), int(auto_reset), _p(obs), _p(mask), _p(reward), _p(done),
                                _p(term), _p(aout))
        out = dict(obs=obs, mask=mask, reward=reward, done=done, actions=aout)
        if want_term:
            out["term_obs"] = term
        return out

    def step_fast(self, auto_reset=True):
        """Fused-sampling step without output buffers (CPU baseline timing)."""
        lib().hexref_batch_step(self._h, None, None, int(auto_reset), None, None, None, None, None, None)

    def set_manual_opponent(self, pool_size=0):
        lib().hexref_batch_set_manual(self._h, pool_size)

    def set_opponent_eps(self, eps):
        lib().hexref_batch_set_opponent_eps(self._h, float(eps))

    def set_eval(self, eval_state):
        lib().hexref_batch_set_eval(self._h, int(bool(eval_state)))

    def half_step(self, side, actions, auto_reset=True, want_term=False):
        reward = np.empty(self.G, np.float32)
        done = np.empty(self.G, np.uint8)
        to_move = np.empty(self.G, np.uint8)
        opp_index = np.empty(self.G, np.int32)
        term = np.zeros((self.G, self.N, self.N), np.int8) if want_term else None
        a = None if actions is None else np.ascontiguousarray(actions, np.int32)
        lib().hexref_batch_half_step(self._h, side, _p(a), int(auto_reset), _p(reward), _p(done), _p(to_move), _p(opp_index), _p(term))
        out = dict(reward=reward, done=done, to_move=to_move, opp_index=opp_index)
        if want_term:
            out["term_obs"] = term
        return out

    def info(self):
        opp = np.empty(self.G, np.int32)
        winner = np.empty(self.G, np.int8)
        lib().hexref_batch_info(self._h, _p(opp), _p(winner))
        return opp, winner

    def opp_state(self):
        to_move = np.empty(self.G, np.uint8)
        opp_index = np.empty(self.G, np.int32)
        lib().hexref_batch_opp_state(self._h, _p(to_move), _p(opp_index))
        return to_move, opp_index

    def view1(self):
        obs = np.empty((self.G, self.N, self.N), np.int8)
        mask = np.empty((self.G, self.C), np.uint8)
        lib().hexref_batch_observe_opponent(self._h, _p(obs), _p(mask))
        return obs, mask

    def observe(self):
        obs = np.empty((self.G, self.N, self.N), np.int8)
        mask = np.empty((self.G, self.C), np.uint8)
        lib().hexref_batch_observe(self._h, _p(obs), _p(mask))
        return obs, mask

    def ply(self, actions):
        ret = np.empty(self.G, np.int8)
        a = np.ascontiguousarray(actions, np.int32)
        lib().hexref_batch_ply(self._h, _p(a), _p(ret))
        return ret

    def env_set_board(self, boards, mask=None):
        """boards in the variant's own codes, true coordinates; BLACK to move."""
        b = np.ascontiguousarray(boards, np.int8)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().hexref_batch_env_set_board(self._h, _p(b), _p(m))

    def set_board(self, boards, cur=0):
        b = np.ascontiguousarray(boards, np.int8)
        lib().hexref_batch_set_board(self._h, _p(b), cur)

    def set_board_labels(self, boards, planes, cur=0):
        """Raw games from preset boards AND their label planes (HexGame.__init__ with connected_stones given)."""
        b = np.ascontiguousarray(boards, np.int8)
        p = np.ascontiguousarray(planes, np.uint8)
        lib().hexref_batch_set_board_labels(self._h, _p(b), _p(p), cur)

    def export(self):
        N, G = self.N, self.G
        out = dict(board=np.empty((G, N, N), np.float64), regions=np.empty((G, 2, N + 2, N + 2), np.float64),
                   region_counter=np.empty((G, 2), np.float64), cur=np.empty(G, np.int8), done=np.empty(G, np.uint8),
                   winner=np.empty(G, np.int8), agent=np.empty(G, np.int8), draws=np.empty(G, np.uint32))
        lib().hexref_batch_export(self._h, *[_p(out[k]) for k in
                                             ("board", "regions", "region_counter", "cur", "done", "winner", "agent", "draws")])
        return out

    def stats(self):
        out = np.zeros(8, np.int64)
        lib().hexref_batch_stats(self._h, _p(out))
        return out


def set_threads(n):
    lib().hexref_set_threads(int(n))


def draw(seed, game, idx):
    return lib().hexref_draw(seed, game, idx)
