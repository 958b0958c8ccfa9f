"""TEST INFRASTRUCTURE ONLY - Python/numpy restatement of the reference's variant-B self-play loop.

Why a second restatement next to oracle/hexref.c: the reference IS a Python/numpy program, and BASELINE.json asks for
"the reference's Python HexGame loop timed on the GPU box's own host cores". /root/reference cannot travel to the GPU box,
so this file restates that loop with the SAME per-step work the reference does (float64 numpy board and padded region
planes, the per-cell Python mask loop, copy + transpose + three masked assignments per perspective switch, whole-plane
`regions[regions == label] = m` relabels) and bench.py times it there (cpu_baseline / --impl reference). It is also
cross-checked against the C oracle in tests/test_pyloop.py, which pins it to the golden vectors transitively.

Follows (paths relative to the reference root):
  Sim.place / Sim.merge_regions   minihex/HexSingleGame.py:88-122 (fast_move), :135-153 (flood_fill), :77-79 (is_valid_move)
  Env.ply / Env.flip              minihex/HexSingleGame.py:233-263 (HexEnv.step), :265-271 (invert_board), :205-206 (legal_actions)
  SelfPlay.*                      minihex/SelfplayWrapper.py:69-89 (reset), :91-104 (setup_opponents), :146-172 (continue_game),
                                  :174-199 (step), :17-22 (BaseRandomPolicy.choose_action)
"""
import os
import random as _random
import time

import numpy as np

OWN, OTHER, FREE = -1.0, 1.0, 0.0


class Sim(object):
    """Board + two padded label planes + counters (HexSingleGame.py:26-71 with an empty board)."""

    def __init__(self, n, template=None):
        self.n = n
        self.board = np.zeros((n, n))
        self.free = n * n
        if template is None:
            planes = np.zeros((2, n + 2, n + 2))
            planes[1][:, 0] = 1
            planes[0][0, :] = 1
            planes[1][:, n + 1] = 2
            planes[0][n + 1, :] = 2
            self.planes = planes
        else:
            self.planes = template.copy()      # cached empty planes (HexSingleGame.py:226-231)
        self.next_label = np.zeros(2)
        self.next_label[0] = np.max(self.planes[0]) + 1
        self.next_label[1] = np.max(self.planes[1]) + 1
        self.to_move = 0
        self.over = False
        self.winner = None

    def cell(self, a):
        r = a // self.n
        return r, a - self.n * r

    def is_free(self, a):
        r, c = self.cell(a)
        return self.board[r, c] == FREE

    def merge_regions(self, r, c):
        plane = self.planes[self.to_move]
        r, c = r + 1, c + 1
        win = plane[r - 1:r + 2, c - 1:c + 2].copy()
        win[0, 0] = 0
        win[2, 2] = 0
        labels = sorted(set(win.flatten().tolist()))
        labels.pop(0)
        if not labels:
            plane[r, c] = self.next_label[self.to_move]
            self.next_label[self.to_move] += 1
        else:
            keep = labels.pop(0)
            plane[r, c] = keep
            for lab in labels:
                plane[plane == lab] = keep

    def place(self, a):
        if not self.is_free(a):
            return 3
        r, c = self.cell(a)
        self.board[r, c] = OWN
        self.free -= 1
        if self.to_move == 1:
            self.merge_regions(c, r)
        else:
            self.merge_regions(r, c)
        won = None
        if self.planes[self.to_move][-1, -1] == 1:
            self.over = True
            won = self.to_move
            self.winner = won
        elif self.free <= 0:
            self.over = True
        self.to_move = (self.to_move + 1) % 2
        return won


class Env(object):
    """One-ply environment (HexSingleGame.py:156-271)."""

    def __init__(self, n):
        self.n = n
        self.template = None
        self.sim = None
        self.mover = 0
        self.last = None

    def reset(self):
        self.mover = 0
        self.sim = Sim(self.n, self.template)
        if self.template is None:
            self.template = self.sim.planes.copy()
        return self.sim.board

    def legal(self):
        return np.array([self.sim.is_free(a) for a in range(self.n * self.n)])

    def flip(self):
        b = self.sim.board.copy()
        t = b.T
        t[t == OWN] = -2
        t[t == OTHER] = OWN
        t[t == -2] = OTHER
        self.sim.board = t

    def ply(self, a):
        self.last = self.sim.place(a)
        if self.last == 3:
            self.sim.over = True
        if self.last == self.mover:
            r = 1
        elif self.last == (self.mover + 1) % 2:
            r = -1
        else:
            r = 0
        rew = [-r, -r]
        rew[self.mover] = r
        self.mover = (self.mover + 1) % 2
        self.flip()
        return self.sim.board, rew, self.sim.over


def random_free_cell(board, rng):
    """BaseRandomPolicy.choose_action (SelfplayWrapper.py:17-22)."""
    idx = np.arange(board.shape[0] * board.shape[1])
    free = idx[board.flatten() == 0]
    return free[int(rng.random() * len(free))]


class SelfPlay(Env):
    """Agent ply + random-opponent reply (SelfplayWrapper.py:37-199, every pool entry a BaseRandomPolicy)."""

    def __init__(self, n, agent=None, rng=None, eval_state=False):
        Env.__init__(self, n)
        self.agent = agent
        self.rng = rng or _random
        self.eval_state = eval_state

    def reset(self):
        Env.reset(self)
        if self.agent is None:
            self.agent = self.rng.randint(0, 1)
        if not self.eval_state:
            if not self.rng.uniform(0, 1) < 0.8:
                self.rng.random()
        if self.mover != self.agent:
            self.reply()
        return self.sim.board

    def reply(self):
        self.rng.uniform(0, 1)                      # unused draw (SelfplayWrapper.py:159)
        self.legal()                                # mask computed for the opponent model, ignored by the random policy (:161)
        a = random_free_cell(self.sim.board, self.rng)
        return self.ply(a)

    def step(self, a):
        obs, rew, over = self.ply(a)
        if not over:
            obs, rew, over = self.reply()
        return obs, rew[self.agent], over


def loop(n, seconds, seed=0):
    """The reference's rollout loop for ONE env: mask = legal_actions(); a = BaseRandomPolicy.choose_action(obs);
    step(a); reset on done. Returns (env steps done, seconds used)."""
    rng = _random.Random(seed)
    env = SelfPlay(n, None, rng)
    obs = env.reset()
    steps, t0 = 0, time.perf_counter()
    while True:
        env.legal()
        a = random_free_cell(obs, rng)
        obs, _, over = env.step(a)
        steps += 1
        if over:
            obs = env.reset()
        if (steps & 63) == 0:
            dt = time.perf_counter() - t0
            if dt >= seconds:
                return steps, dt


def _worker(args):
    n, seconds, seed = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    return loop(n, seconds, seed)


def rate(n, seconds, procs):
    """Aggregate env-steps/s of `procs` independent single-env loops (one process per core), plus a description."""
    if procs <= 1:
        s, dt = loop(n, seconds, 0)
        return s / dt, "1 process x %.1f s, %d env steps (python/numpy restatement of the minihex loop)" % (dt, s)
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        res = pool.map(_worker, [(n, seconds, i) for i in range(procs)])
    total = sum(s / dt for s, dt in res)
    return total, "%d processes x %.1f s, %d env steps in total (python/numpy restatement of the minihex loop)" % (
        procs, seconds, sum(s for s, _ in res))


if __name__ == "__main__":
    import json
    import sys
    _n, _sec, _procs = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3])
    _v, _s = rate(_n, _sec, _procs)
    print(json.dumps({"value": _v, "sample": _s, "procs": _procs}))
