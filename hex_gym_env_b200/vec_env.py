"""HexVecEnv - G games as ONE vectorised environment, shaped like a stable-baselines3 `VecEnv`.

This is what replaces `DummyVecEnv([lambda: ActionMasker(SelfPlayEnv(...), mask_fn)])` of the reference training scripts
(scripts/experiments/*.py:34-47) when the rollout is collected from many games at once: one fused kernel launch per
`step()` for all games instead of one Python env per process.

SB3 surface mirrored (stable-baselines3 2.2.1 / sb3-contrib, from memory - SB3 is not installed in the build image, so
this contract is duck-typed; if gymnasium / SB3 are importable the spaces are real gymnasium spaces):
  num_envs, observation_space, action_space
  reset() -> obs[G,N,N]
  step_async(actions); step_wait() -> (obs, rewards, dones, infos);  step(actions)
  auto-reset on done, with infos[i]["terminal_observation"] for the finished games
  env_method("action_masks") / action_masks() -> bool[G,C]      (what sb3_contrib's get_action_masks() calls)
  get_attr / set_attr / env_is_wrapped / seed / close
Learned opponents: base_model= / buffer_size= / scores= as in selfplay_wrapper(HexEnv)(...) give every game the reference's opponent
buffer (one OpponentPool, opponents.py); its API (set_eval, get_scores, set_opponent_model, ...) is reachable on the env and
through env_method, which is what SelfPlayCallback (minihex/EvaluationCallback.py) calls on its gym_env.
variant="hex-v0" with opponent_model= / eps=: gym.make("hex-v0", opponent_policy="opponent_predict", opponent_model=..., eps=...)
(scripts/selfplay.py:38-44) for every game - a batched policy on the opponent's view, random_policy with probability eps.

`output="numpy"` (default, what SB3 expects) copies results to host arrays; `output="torch"` returns device tensors and
never touches the host (use it when the policy lives on the same GPU; `infos` is then a lazy object, not a list).
"""
import numpy as np
import torch

from .batch import AGENT_BLACK, AGENT_RANDOM, AGENT_WHITE, VARIANT_A, VARIANT_B, HexBatch
from .minihex_compat import _spaces


class _LazyInfos(object):
    """List-of-dicts view over the step's device results, materialised only for the indices that are read."""

    def __init__(self, done, term_obs, n):
        self._done, self._term, self._n = done, term_obs, n

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if bool(self._done[i]):
            t = self._term[i]
            return {"terminal_observation": t if isinstance(t, np.ndarray) else t.clone(), "TimeLimit.truncated": False}
        return {}

    def __iter__(self):
        return (self[i] for i in range(self._n))


def random_start_boards(num_games, board_size, generator, device):
    """Batched stand-in for HexEnv.random_board (HexSingleGame.py:300-331): per game a random rectangle (sides in
    [N//4, N-2], uniform position) of the empty board receives an even number of stones, about 50-100 % of its cells, half
    BLACK and half WHITE, at random cells of the rectangle - so BLACK is to move. The reference draws from the global
    np.random; this generator is torch's (keyed by `generator`), i.e. the same distribution family, not the same numbers.
    Returns int8[G,N,N] in true coordinates: 0 BLACK, 1 WHITE, 2 EMPTY."""
    G, N = num_games, board_size
    lo, hi = N // 4, max(N - 1, N // 4 + 1)
    r = lambda a, b: torch.randint(a, b, (G,), generator=generator, device=device)
    rows, cols = r(lo, hi), r(lo, hi)
    top = (torch.rand(G, generator=generator, device=device) * (N - rows + 1).float()).long()
    left = (torch.rand(G, generator=generator, device=device) * (N - cols + 1).float()).long()
    frac = 0.5 + 0.5 * torch.rand(G, generator=generator, device=device)
    stones = ((rows * cols).float() * frac / 2).floor().long() * 2
    yy = torch.arange(N, device=device).view(1, N, 1)
    xx = torch.arange(N, device=device).view(1, 1, N)
    inside = (yy >= top.view(G, 1, 1)) & (yy < (top + rows).view(G, 1, 1)) & (xx >= left.view(G, 1, 1)) & (xx < (left + cols).view(G, 1, 1))
    keys = torch.rand((G, N, N), generator=generator, device=device).masked_fill(~inside, 2.0).view(G, -1)
    rank = keys.argsort(dim=1).argsort(dim=1)
    board = torch.full((G, N * N), 2, dtype=torch.int8, device=device)
    board[rank < (stones // 2).view(G, 1)] = 0
    board[(rank >= (stones // 2).view(G, 1)) & (rank < stones.view(G, 1))] = 1
    return board.view(G, N, N)


try:  # subclass the real VecEnv when stable-baselines3 is installed, so that isinstance checks inside SB3 pass
    from stable_baselines3.common.vec_env.base_vec_env import VecEnv as _VecEnvBase
except Exception:  # pragma: no cover - SB3 is absent in the build image
    _VecEnvBase = object


class HexVecEnv(_VecEnvBase):
    def __init__(self, board_size=5, num_envs=1024, variant="selfplay", agent_player_num=None, opponent_first=False,
                 seed=0, device=None, output="numpy", obs_dtype=None, game_offset=0, sample_board=False,
                 base_model=None, buffer_size=20, scores=None, opponent_model=None, eps=0.5, info_fields=False):
        if variant in ("selfplay", "B", VARIANT_B):
            v = VARIANT_B
            agent_mode = AGENT_RANDOM if agent_player_num is None else (AGENT_WHITE if int(agent_player_num) else AGENT_BLACK)
            low, high = -1, 1
        elif variant in ("hex-v0", "A", VARIANT_A):
            v, agent_mode, low, high = VARIANT_A, AGENT_BLACK, 0, 2
        else:
            raise ValueError("variant must be 'selfplay' (SelfPlayEnv, variant B) or 'hex-v0' (HexEnv, variant A)")
        if output not in ("numpy", "torch"):
            raise ValueError("output must be 'numpy' or 'torch'")
        self.num_envs, self.board_size, self.output = int(num_envs), int(board_size), output
        self.sample_board = bool(sample_board)
        if sample_board and v != VARIANT_B:
            raise ValueError("sample_board is a SelfPlayEnv (variant B) option (HexSingleGame.py:171)")
        # sample_board: episodes start from random positions. The fused step kernel restarts games on the empty board, so this
        # mode runs the split step instead (agent ply / built-in random opponent ply as separate launches around the import).
        # base_model: learned opponents - selfplay_wrapper(HexEnv)(base_model=..., scores=..., buffer_size=...) for every game
        # (SelfplayWrapper.py:39-67): a batched policy (see opponents.py) fills the pool; the split step plays its entries.
        if base_model is not None and (v != VARIANT_B or sample_board):
            raise ValueError("base_model (an opponent pool) belongs to the SelfPlayEnv variant without sample_board")
        # opponent_model / eps: gym.make("hex-v0", opponent_policy="opponent_predict", opponent_model=..., eps=...) for every game
        # (HexGame.py:165-167,354-359; scripts/selfplay.py:38-44): a batched policy answers on the opponent's view, and with
        # probability eps - drawn from the game's own stream on the device - random_policy moves instead.
        if opponent_model is not None and v != VARIANT_A:
            raise ValueError("opponent_model / eps (opponent_predict) belong to the 'hex-v0' variant; SelfPlayEnv takes base_model")
        self.batch = HexBatch(board_size, num_envs, variant=v, device=device, seed=seed, game_offset=game_offset,
                              agent_mode=agent_mode, opponent_first=opponent_first, auto_reset=True,
                              manual_opponent=self.sample_board or base_model is not None or opponent_model is not None,
                              pool_size=int(buffer_size) if base_model is not None else 0)
        self.pool = None
        if base_model is not None:
            from .opponents import OpponentPool
            self.pool = OpponentPool(base_model, buffer_size=int(buffer_size), scores=scores, batch=self.batch)
        # info_fields: the info dict of variant-A HexEnv.step (HexGame.py:281-286) per env: last_move_opponent, last_move_player, winner
        self.info_fields = bool(info_fields)
        if self.info_fields:
            if v != VARIANT_A or opponent_model is not None:
                raise ValueError("info_fields is the info dict of the fused 'hex-v0' step (HexGame.py:281-286)")
            self.batch.enable_info()
        self.opponent_model, self.eps = opponent_model, (float(eps) if opponent_model is not None else None)
        if opponent_model is not None:
            self.batch.set_opponent_eps(self.eps)
        self.device = self.batch.device
        self._bgen = torch.Generator(device=self.device)
        self._bgen.manual_seed(int(seed) + 0x5EED)
        self.obs_dtype = obs_dtype or (np.float32 if output == "numpy" else torch.float32)
        self.observation_space = _spaces.Box(low=low, high=high, shape=(board_size, board_size),
                                             dtype=np.int64 if v == VARIANT_B else np.uint8)
        self.action_space = _spaces.Discrete(board_size ** 2)
        if _VecEnvBase is not object:
            _VecEnvBase.__init__(self, self.num_envs, self.observation_space, self.action_space)
        self.render_mode = None
        # SB3 >= 2.0 VecEnv bookkeeping (base_vec_env.py: __init__): per-env reset infos, seeds and options of the next reset
        self.reset_infos = [{} for _ in range(self.num_envs)]
        self._seeds = [None for _ in range(self.num_envs)]
        self._options = [{} for _ in range(self.num_envs)]
        self._actions = None
        self._mask = None
        self._pinned = None

    # ------------------------------------------------------------------ helpers
    def _host(self, name, t):
        """Device tensor -> numpy through a reusable pinned staging buffer."""
        if self._pinned is None:
            self._pinned = {}
        p = self._pinned.get(name)
        if p is None or p.shape != t.shape or p.dtype != t.dtype:
            p = torch.empty(t.shape, dtype=t.dtype).pin_memory()
            self._pinned[name] = p
        p.copy_(t, non_blocking=True)
        return p

    def _obs_out(self, obs):
        if self.output == "torch":
            return obs.to(self.obs_dtype)
        h = self._host("obs", obs)
        torch.cuda.current_stream(self.device).synchronize()
        return h.numpy().astype(self.obs_dtype)

    # ------------------------------------------------------------------ VecEnv API
    def _restart_on_sampled_boards(self, which=None):
        boards = random_start_boards(self.num_envs, self.board_size, self._bgen, self.device)
        self.batch.import_boards(boards, import_mask=which)

    def reset(self):
        self.reset_infos = [{} for _ in range(self.num_envs)]
        self._reset_seeds()
        self._reset_options()
        obs, mask = self.batch.reset()
        if self.sample_board:
            self._restart_on_sampled_boards()
            self.batch.opponent_catch_up()      # SelfPlayEnv.reset -> continue_game where the opponent moves first
            obs, mask = self.batch.encode(0)
        elif self._opponent_fn() is not None:
            self.batch.opponent_opening(self._opponent_fn())   # the opponent opens where it moves first (:79-80, HexGame.py:224-230)
            obs, mask = self.batch.encode(0)
        self._mask = mask
        return self._obs_out(obs)

    def _step_sampled(self, actions):
        """One env step in sample_board mode, in the reference's order: agent ply -> (finished: new sampled board) ->
        opponent reply / opening -> (finished: new sampled board) -> opening."""
        b = self.batch
        term = b._buf("sb_term", (b.G, b.N, b.N), b.obs_dtype)
        h = b.half_step(0, actions, term_obs=term)
        reward, done = h["reward"].clone(), h["done"].clone()
        self._restart_on_sampled_boards(h["done"])
        h = b.half_step(1, None, term_obs=term)
        reward += h["reward"]
        done |= h["done"]
        self._restart_on_sampled_boards(h["done"])
        b.opponent_catch_up()
        obs, mask = b.encode(0)
        return dict(obs=obs, mask=mask, reward=reward, done=done, term_obs=term)

    def step_async(self, actions):
        if isinstance(actions, np.ndarray):
            actions = torch.from_numpy(np.ascontiguousarray(actions.astype(np.int32, copy=False)))
        self._actions = actions

    def step_wait(self):
        if self.sample_board:
            o = self._step_sampled(self._actions)
        elif self._opponent_fn() is not None:
            o = self.batch.step_with_opponent(self._actions, self._opponent_fn(), want_term=True)
        else:
            o = self.batch.step(self._actions, want_term=True, want_actions=self.info_fields)
        self._mask = o["mask"]
        if self.output == "torch":
            infos = _LazyInfos(o["done"], o["term_obs"], self.num_envs)
            if self.info_fields:   # device tensors, -1 = None (winner 3 = illegal move), overwritten by the next step
                infos.last_move_opponent, infos.winner, infos.last_move_player = self.batch.last_move_opponent, self.batch.winner, o["actions"]
            return o["obs"].to(self.obs_dtype), o["reward"], o["done"].bool(), infos
        keys = ("obs", "reward", "done", "term_obs")
        h = {k: self._host(k, o[k]) for k in keys}
        if self.info_fields:
            h.update(lmo=self._host("lmo", self.batch.last_move_opponent), win=self._host("win", self.batch.winner),
                     lmp=self._host("lmp", o["actions"]))
        torch.cuda.current_stream(self.device).synchronize()
        done = h["done"].numpy().astype(bool)
        term = h["term_obs"].numpy()
        if self.info_fields:   # HexGame.py:281-286 ('state' is the observation itself)
            lmo, win, lmp = h["lmo"].numpy(), h["win"].numpy(), h["lmp"].numpy()
            infos = [{"last_move_opponent": None if lmo[i] < 0 else int(lmo[i]), "last_move_player": int(lmp[i]),
                      "winner": None if win[i] < 0 else int(win[i])} for i in range(self.num_envs)]
        else:
            infos = [{} for _ in range(self.num_envs)]
        for i in np.flatnonzero(done):
            infos[i].update({"terminal_observation": term[i].astype(self.obs_dtype), "TimeLimit.truncated": False})
        return h["obs"].numpy().astype(self.obs_dtype), h["reward"].numpy().copy(), done, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def action_masks(self):
        """bool[G,C] legal-action masks of the current observations (HexEnv.legal_actions / get_action_mask)."""
        if self._mask is None:
            _, self._mask = self.batch.encode(0)
        if self.output == "torch":
            return self._mask.bool()
        h = self._host("mask", self._mask)
        torch.cuda.current_stream(self.device).synchronize()
        return h.numpy().astype(bool)

    # the opponent buffer of SelfPlayEnv (SelfplayWrapper.py:106-144), one pool for all games: what SelfPlayCallback calls on its
    # gym_env (EvaluationCallback.py:31-50). Also reachable through env_method, which answers once per env like SB3 does.
    _POOL_API = ("set_eval", "get_scores", "get_opponent_models", "set_opponent_model", "append_opponent_model",
                 "get_best_mean_reward", "save_best_model")
    _POOL_ATTRS = ("opponent_models", "opponent_scores", "best_model", "best_score", "best_mean_reward", "eval_state")

    def _opponent_fn(self):
        """The caller-driven opponent of the split step: the pool (variant B) or the one opponent_predict model (variant A)."""
        if self.pool is not None:
            return self.pool
        if self.opponent_model is not None:
            return self._single_opponent
        return None

    def _single_opponent(self, obs, mask, to_move, opp_index):
        return self.opponent_model(obs, mask)

    def set_opponent_model(self, *args, **kwargs):
        """SelfPlayEnv.set_opponent_model(index, model, score) (SelfplayWrapper.py:125-136) with a pool; variant A:
        HexEnv.set_opponent_model(model) (HexGame.py:351-352)."""
        if self.pool is not None:
            return self.pool.set_opponent_model(*args, **kwargs)
        if self.opponent_model is None:
            raise AttributeError("set_opponent_model needs learned opponents: create the HexVecEnv with base_model= or opponent_model=")
        (self.opponent_model,) = args

    def __getattr__(self, name):
        if name in HexVecEnv._POOL_API or name in HexVecEnv._POOL_ATTRS:
            pool = self.__dict__.get("pool")
            if pool is None:
                raise AttributeError("%r needs an opponent pool: create the HexVecEnv with base_model=..." % (name,))
            return getattr(pool, name)
        raise AttributeError(name)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        idx = range(self.num_envs) if indices is None else ([indices] if isinstance(indices, int) else indices)
        if method_name in ("action_masks", "legal_actions", "get_action_mask"):
            m = self.action_masks()
            return [m[i] for i in idx]
        if method_name == "set_opponent_model" and (self.pool is not None or self.opponent_model is not None):
            r = self.set_opponent_model(*args, **kwargs)
            return [r for _ in idx]
        if method_name in HexVecEnv._POOL_API and self.pool is not None:
            r = getattr(self.pool, method_name)(*args, **kwargs)     # one pool for all games: called once
            return [r for _ in idx]
        raise AttributeError("HexVecEnv has no per-env method %r" % (method_name,))

    def get_attr(self, attr_name, indices=None):
        n = self.num_envs if indices is None else (1 if isinstance(indices, int) else len(indices))
        return [getattr(self, attr_name)] * n

    def set_attr(self, attr_name, value, indices=None):
        setattr(self, attr_name, value)

    def env_is_wrapped(self, wrapper_class, indices=None):
        n = self.num_envs if indices is None else (1 if isinstance(indices, int) else len(indices))
        return [False] * n

    def seed(self, seed=None):
        """SB3: list of the seeds handed to the envs. Recorded, not used: the reference accepts and ignores reset(seed=...)
        (HexGame.py:206, HexSingleGame.py:208, SelfplayWrapper.py:69); games are keyed by the constructor's seed."""
        self._seeds = [None if seed is None else seed + i for i in range(self.num_envs)]
        return list(self._seeds)

    def set_options(self, options=None):
        if options is None:
            options = {}
        self._options = [dict(options) for _ in range(self.num_envs)] if isinstance(options, dict) else [dict(o) for o in options]

    def _reset_seeds(self):
        self._seeds = [None for _ in range(self.num_envs)]

    def _reset_options(self):
        self._options = [{} for _ in range(self.num_envs)]

    @property
    def unwrapped(self):
        return self

    def getattr_depth_check(self, name, already_found):
        return None

    def get_images(self):
        return [None for _ in range(self.num_envs)]

    def render(self, mode=None):
        return None

    def episode_stats(self):
        from .batch import STAT_NAMES
        return dict(zip(STAT_NAMES, self.batch.stats().cpu().tolist()))

    def close(self):
        self.batch.close()
