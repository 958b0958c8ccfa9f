"""Single-environment drop-ins for the minihex classes, with the game core on the GPU.

`from hex_gym_env_b200 import minihex` gives the names the reference scripts import (HexGame, HexEnv, player,
selfplay_wrapper, BaseRandomPolicy, OpponentPolicy, random_policy) with the reference's constructor arguments, return
conventions and error behaviour, so `scripts/experiments/*.py` and `scripts/selfplay*.py` of MBPrdctns/hex_gym_env run
against it unchanged. Every stone placement, region-label merge, win check, validity test and board view goes through
libhexb.so (hexb_ply / hexb_export_state / hexb_encode / hexb_import_boards on a one-game raw handle); what stays on the
host is what the reference also keeps outside HexGame: the env-level reward / done rules, the opponent-pool bookkeeping
and the calls into Python's global `random`, made in exactly the reference's order so that a seeded reference run and a
seeded run of these classes see the same draws.

`selfplay_wrapper` below restates the bodies of minihex/SelfplayWrapper.py:37-199 (GUI branches and prints removed) on purpose:
its method names, attribute names (opponent_models, opponent_scores, best_model, eval_state, ...) and the order of its calls
into `random` ARE the contract the reference's callbacks and scripts rely on ("pool bookkeeping stays
host-side Python"), so that class is a line-by-line host-side mirror, not new design; the game core underneath it is the GPU's.

For throughput use the batched classes (HexBatch, vec_env.HexVecEnv): these single-game facades pay one kernel launch and
one device->host read per call and exist for compatibility and for parity tests that read like the reference's own usage.

Reference lines (paths relative to the reference root):
  DeviceGame             minihex/HexGame.py:16-142 (variant A), minihex/HexSingleGame.py:21-153 (variant B)
  HexEnvA                minihex/HexGame.py:145-371
  HexEnvB                minihex/HexSingleGame.py:156-331
  selfplay_wrapper       minihex/SelfplayWrapper.py:37-208
  BaseRandomPolicy / OpponentPolicy / random_policy   SelfplayWrapper.py:16-35, minihex/__init__.py:8-12
Not carried over (out of scope): the pygame GUI (play_gui / show_board / "interactive"),
player_color=WHITE of variant A (corrupt in the reference, HexGame.py:245-248).
"""
import random  # noqa: F401  (module attribute on purpose: the reference draws from the global `random`; tests swap it)
from enum import IntEnum

import numpy as np

from .batch import VARIANT_A, VARIANT_B, HexBatch

try:  # the real gymnasium when it is installed (SB3 checks isinstance(env, gymnasium.Env)); a stand-in otherwise
    import gymnasium as _gym
    from gymnasium import spaces as _spaces
    _EnvBase = _gym.Env
except Exception:  # pragma: no cover - gymnasium is absent in the build image
    _gym = None
    _EnvBase = object

    class _Space(object):
        def __init__(self, **kw):
            self.__dict__.update(kw)

    class _spaces(object):
        @staticmethod
        def Box(low, high, shape, dtype):
            return _Space(low=low, high=high, shape=shape, dtype=dtype)

        @staticmethod
        def Discrete(n):
            return _Space(n=n)


class player(IntEnum):
    """Variant-A cell / player codes (HexGame.py:10-13)."""
    BLACK = 0
    WHITE = 1
    EMPTY = 2


# Variant-B codes (HexSingleGame.py:15-19, SelfplayWrapper.py:10-14)
player_b = {
    "BLACK": {"id": 0, "board_encoding": -1},
    "WHITE": {"id": 1, "board_encoding": 1},
    "EMPTY": {"id": 2, "board_encoding": 0},
}

_DEFAULT_DEVICE = [None]
_HANDLE_POOL = {}   # (variant, board size, device) -> idle one-game device handles, reused across episodes


def set_device(device):
    """CUDA device ordinal used by the single-game facades (default: torch's current device)."""
    _DEFAULT_DEVICE[0] = device


def _swap_view(board, lo, hi):
    """invert_board (HexGame.py:297-303 / HexSingleGame.py:265-271): transpose and exchange the two stone codes."""
    t = np.array(board).T.copy()
    a, b = t == lo, t == hi
    t[a], t[b] = hi, lo
    return t


# ---------------------------------------------------------------------------------------------------- game core
class DeviceGame(object):
    """One Hex position held in GPU memory. variant A: HexGame(current_player_num, board, focus_player, ...);
    variant B: HexGame(active_player, board, ...). See HexGameA / HexGameB below for the reference signatures."""

    _variant = VARIANT_A
    _empty = 2

    def _setup(self, to_move, board, connected_stones, debug):
        board = np.asarray(board)
        n = board.shape[1]
        self._n = n
        self._pool_key = (self._variant, n, _DEFAULT_DEVICE[0])
        idle = _HANDLE_POOL.setdefault(self._pool_key, [])
        self._dev = None
        while idle and self._dev is None:
            cand = idle.pop()
            self._dev = cand if getattr(cand, "_h", None) else None   # never reuse a handle the garbage collector closed
        if self._dev is None:
            self._dev = HexBatch(n, 1, variant=self._variant, device=_DEFAULT_DEVICE[0], raw=True)
        self._dev.reset()
        if self._variant == VARIANT_A:
            true_codes = np.where(board == 0, 0, np.where(board == 1, 1, 2)).astype(np.int8)
        else:
            true_codes = np.where(board == -1, 0, np.where(board == 1, 1, 2)).astype(np.int8)
        stones = int((true_codes != 2).sum())
        if connected_stones is not None and stones:
            # the planes are adopted as they are and region_counter = max(plane) + 1 (HexGame.py:46-51 / HexSingleGame.py:50-55):
            # HexEnv.reset from its second call on (cached planes) and user-supplied `regions=`. After a merge the highest
            # label can be lower than a raster-order rebuild's counter, so this is not the same state as the branch below.
            planes = np.ascontiguousarray(np.asarray(connected_stones), dtype=np.uint8)
            self._dev.import_labels(true_codes[None], planes[None], np.array([int(to_move)], np.int8))
        elif stones or int(to_move) != 0:
            # preset position: stones enter in raster order through flood_fill (HexGame.py:53-61 / HexSingleGame.py:57-65)
            self._dev.import_boards(true_codes[None], np.array([int(to_move)], np.int8))
        self.empty_fields = int(np.count_nonzero(board == self._empty))
        self._done_host = False
        self._ref_flipped = bool(int(to_move)) if self._variant == VARIANT_B else False  # parity of the env's invert_board() calls
        self.debug = debug   # (make_move dispatches on it; no bound method is stored on the instance: that would be a reference cycle)
        self.actions = np.arange(n * n)
        self._cache = None

    def __del__(self):
        dev, key = getattr(self, "_dev", None), getattr(self, "_pool_key", None)
        if dev is not None and key is not None and _HANDLE_POOL is not None and getattr(dev, "_h", None):
            self._dev = None
            if len(_HANDLE_POOL.setdefault(key, [])) < 8:
                _HANDLE_POOL[key].append(dev)

    # -- state, read back from the device in the reference's layout (float64 like the reference's numpy arrays)
    def _state(self):
        if self._cache is None:
            self._cache = {k: v.copy() for k, v in self._dev.export_state_host().items()}   # one kernel + one copy
        return self._cache

    @property
    def board_size(self):
        return self._n

    @property
    def board(self):
        st = self._state()
        b = st["board"][0]
        if self._variant == VARIANT_B and self._ref_flipped != bool(st["cur"][0]):
            # the device shows the board from the side to move; the reference shows it after as many invert_board() calls
            # as the env made. They only differ after an illegal move (the env flips although nobody moved).
            b = _swap_view(b, -1.0, 1.0)
        return b

    @property
    def regions(self):
        return self._state()["regions"][0]

    @property
    def region_counter(self):
        return self._state()["region_counter"][0]

    @property
    def current_player_num(self):
        return int(self._state()["cur"][0])

    @property
    def winner(self):
        w = int(self._state()["winner"][0])
        return None if w < 0 else (player(w) if self._variant == VARIANT_A else w)

    @property
    def done(self):
        return bool(self._done_host or self._state()["done"][0])

    @done.setter
    def done(self, value):  # the envs write simulator.done = True after an illegal move
        self._done_host = bool(value)

    # -- moves
    def action_to_coordinate(self, action):
        y = action // self._n
        return (y, action - self._n * y)

    def coordinate_to_action(self, coords):
        return np.ravel_multi_index(coords, (self._n, self._n))

    def is_valid_move(self, action):
        y, x = self.action_to_coordinate(action)
        return self.board[y, x] == self._empty  # IndexError for an out-of-range action, like the reference

    def get_possible_actions(self):
        return self.actions[self.board.flatten() == self._empty]

    def make_move(self, action):
        """HexGame.make_move: fast_move, or make_move_debug when the game was built with debug=True (HexGame.py:29-32)."""
        return self.make_move_debug(action) if self.debug else self.fast_move(action)

    def make_move_debug(self, action):
        if not self.is_valid_move(action):
            raise IndexError("Illegal move %s" % (self.action_to_coordinate(action),))
        return self.fast_move(action)

    def fast_move(self, action):
        """Place a stone for the side to move. Returns None, the winner, or 3 for an illegal move (state untouched)."""
        y, x = self.action_to_coordinate(int(action))
        if not (0 <= int(action) < self._n * self._n):
            raise IndexError("index %d is out of bounds for a %dx%d board" % (int(action), self._n, self._n))
        st = self._dev.export_state_host(ply_actions=[int(action)])   # move + fresh state in one round trip
        ret = int(st["ret"][0])
        if ret == 3:
            return 3
        self._cache = {k: v.copy() for k, v in st.items()}
        self.empty_fields -= x if self._variant == VARIANT_A else 1  # variant A: sic, HexGame.py:96
        if ret < 0:
            return None
        return player(ret) if self._variant == VARIANT_A else ret


class HexGameA(DeviceGame):
    """minihex.HexGame.HexGame (variant A: board 0/1/2 in true coordinates)."""
    _variant, _empty = VARIANT_A, 2

    def __init__(self, current_player_num, board, focus_player, connected_stones=None, debug=False):
        self.player = focus_player
        self._setup(current_player_num, board, connected_stones, debug)


class HexGameB(DeviceGame):
    """minihex.HexSingleGame.HexGame (variant B: board -1/+1/0 in the perspective of the side to move, as HexEnv keeps it)."""
    _variant, _empty = VARIANT_B, 0

    def __init__(self, active_player, board, connected_stones=None, debug=False):
        self._setup(active_player, board, connected_stones, debug)


# ---------------------------------------------------------------------------------------------------- policies
def random_policy(board):
    """minihex/__init__.py:8-12 - uniformly random empty cell (code 2) of a variant-A board."""
    cells = np.arange(board.shape[0] * board.shape[1])
    free = cells[board.flatten() == 2]
    return free[int(random.random() * len(free))]


class BaseRandomPolicy(object):
    """SelfplayWrapper.py:16-24 - uniformly random empty cell (code 0) of a variant-B board."""

    def choose_action(self, board, action_mask=None):
        cells = np.arange(board.shape[0] * board.shape[1])
        free = cells[board.flatten() == 0]
        return free[int(random.random() * len(free))]

    def save_model(self, path):
        return None


class OpponentPolicy(object):
    """SelfplayWrapper.py:26-35 - a stable-baselines model as the opponent."""

    def __init__(self, model):
        self.opponent_model = model

    def choose_action(self, board, action_mask=None):
        action, _ = self.opponent_model.predict(board, deterministic=False, action_masks=action_mask)
        return action

    def save_model(self, path):
        self.opponent_model.save(path)


# ---------------------------------------------------------------------------------------------------- variant-A env
def _print_board(board, empty, black):
    """ANSI rendering in the reference's rhombus layout (HexGame.py:305-330 / HexSingleGame.py:273-298): O empty, B / W stones."""
    n = board.shape[1]
    print(" " * 6 + "".join("  %d  |" % (j + 1) for j in range(n)))
    print(" " * 5 + "-" * (n * 6 - 1))
    for i in range(n):
        cells = "".join("  %s  |" % ("O" if board[i, j] == empty else ("B" if board[i, j] == black else "W")) for j in range(n))
        print(" " * (1 + i * 3) + "%d  |" % (i + 1) + cells)
        print(" " * (i * 3 + 1) + "-" * (n * 7 - 1))


class HexEnvA(_EnvBase):
    """minihex.HexGame.HexEnv ('hex-v0'): the agent plays BLACK against `opponent_policy` inside step()."""

    metadata = {"render.modes": ["ansi"]}

    def __init__(self, opponent_policy, opponent_model=None, player_color=player.BLACK, current_player_num=player.BLACK,
                 board=None, regions=None, board_size=5, debug=False, show_board=False, eps=0.5):
        if opponent_policy == "interactive" or show_board:
            raise NotImplementedError("the pygame GUI of the reference is out of scope")
        if int(player_color) != int(player.BLACK):
            raise ValueError("player_color=WHITE corrupts board/regions in the reference (HexGame.py:245-248) and is rejected")
        self.opponent_policy = self.opponent_predict if opponent_policy == "opponent_predict" else opponent_policy
        self.interactive = False
        if board is None:
            board = player.EMPTY * np.ones((board_size, board_size))
        self.n_players = 2
        self.eps = eps
        self.opponent_model = opponent_model
        self.initial_board = board
        self.current_player_num = current_player_num
        self.player = player_color
        self.simulator = None
        self.winner = None
        self.previous_opponent_move = None
        self.debug = debug
        self.board_size = board_size
        self.observation_space = _spaces.Box(low=0, high=2, shape=(board_size, board_size), dtype=np.uint8)
        self.action_space = _spaces.Discrete(board_size ** 2)
        self.initial_regions = regions

    @property
    def opponent(self):
        return player((self.player + 1) % 2)

    def get_action_mask(self):
        view = getattr(self, "_opponent_view", None)
        board = self.simulator.board if view is None else view   # (opponent_predict asks for the mask of the inverted board, HexGame.py:358)
        return board.flatten() == player.EMPTY

    def reset(self, seed=None, options=None):
        cached = self.initial_regions      # None at the first reset (planes rebuilt), adopted afterwards (HexGame.py:207-220)
        self.simulator = HexGameA(self.current_player_num, np.array(self.initial_board), self.player,
                                  connected_stones=cached, debug=self.debug)
        if cached is None:
            self.initial_regions = self.simulator.regions.copy()
        self.previous_opponent_move = None
        if self.player != self.current_player_num:
            self.opponent_move(None)
        info = {"state": self.simulator.board, "last_move_opponent": self.previous_opponent_move, "last_move_player": None}
        return self.simulator.board, info

    def opponent_move(self, info):
        # the opponent sees the transposed, colour-swapped board and its action is transposed back (HexGame.py:332-346)
        seen = _swap_view(self.simulator.board, player.BLACK, player.WHITE)
        self._opponent_view = seen         # while the opponent chooses, the reference's simulator.board IS this inverted board
        try:
            a = self.opponent_policy(seen)
        finally:
            self._opponent_view = None
        y, x = self.simulator.action_to_coordinate(a)
        a = self.simulator.coordinate_to_action((x, y))
        self.winner = self.simulator.make_move(a)
        self.previous_opponent_move = a
        return a

    def step(self, action):
        if not self.simulator.done:
            self.winner = self.simulator.make_move(action)
            if self.winner == 3:
                self.simulator.done = True
        opponent_action = None
        if not self.simulator.done:
            opponent_action = self.opponent_move(None)
        if self.winner == self.player:
            reward = 1
        elif self.winner == self.opponent:
            reward = -1
        elif self.winner == 3:
            reward = -100
        else:
            reward = 0
        info = {"state": self.simulator.board, "last_move_opponent": opponent_action, "last_move_player": action,
                "winner": self.winner}
        return self.simulator.board, reward, self.simulator.done, False, info

    def set_opponent_model(self, model):
        self.opponent_model = model

    def opponent_predict(self, state):
        rv = random.uniform(0, 1)
        if rv < self.eps:
            return random_policy(state)
        action, _ = self.opponent_model.predict(state, deterministic=True, action_masks=self.get_action_mask())
        return action

    def render(self, mode="ansi", close=False):
        _print_board(self.simulator.board, player.EMPTY, player.BLACK)


# ---------------------------------------------------------------------------------------------------- variant-B env
def random_board(matrix):
    """HexEnv.random_board (HexSingleGame.py:300-331): a random rectangle of the empty board filled with equally many
    stones of both colours (an even count, so BLACK is to move). Same np.random call order as the reference."""
    n = matrix.shape[0]
    rows = np.random.randint(n // 4, n - 1)
    cols = np.random.randint(n // 4, n - 1)
    top = np.random.randint(0, n - rows + 1)
    left = np.random.randint(0, n - cols + 1)
    cells = rows * cols
    stones = int((rows * cols * (0.5 + 0.5 * np.random.random())) // 2) * 2
    blacks = stones // 2
    whites = stones - blacks
    values = np.array([-1] * blacks + [1] * whites + [0] * (cells - stones))
    np.random.shuffle(values)
    matrix[top:top + rows, left:left + cols] = values.reshape((rows, cols))
    return matrix


class HexEnvB(_EnvBase):
    """minihex.HexSingleGame.HexEnv: one ply per step(), observation in the perspective of the side to move."""

    metadata = {"render.modes": ["ansi"]}

    def __init__(self, current_player_num=0, board=None, regions=None, board_size=5, debug=False, show_board=False, eps=0.5,
                 sample_board=False):
        if show_board:
            raise NotImplementedError("the pygame GUI of the reference is out of scope")
        if board is None and not sample_board:
            board = np.zeros((board_size, board_size))
        elif sample_board:
            board = random_board(np.zeros((board_size, board_size)))
        self.sample_board = sample_board
        self.eps = eps
        self.initial_board = board
        self.current_player_num = current_player_num
        self.simulator = None
        self.winner = None
        self.debug = debug
        self.board_size = board_size
        self.observation_space = _spaces.Box(low=-1, high=1, shape=(board_size, board_size), dtype=int)
        self.action_space = _spaces.Discrete(board_size ** 2)
        self.initial_regions = regions

    @property
    def observation(self):
        return self.simulator.board

    def legal_actions(self):
        return self.simulator.board.flatten() == 0

    def reset(self, seed=None, options=None):
        self.current_player_num = 0
        if self.sample_board:
            start = random_board(np.zeros((self.board_size, self.board_size)))
        else:
            start = np.array(self.initial_board)
        # sample_board: a fresh position every time, its planes rebuilt (HexSingleGame.py:217-222); otherwise the planes cached at
        # the first reset - or handed in as regions= - are adopted from then on (:211-216, :223-229)
        cached = None if self.sample_board else self.initial_regions
        self.simulator = HexGameB(self.current_player_num, start, connected_stones=cached, debug=self.debug)
        if cached is None:
            self.initial_regions = self.simulator.regions.copy()
        return self.simulator.board

    def step(self, action):
        self.winner = self.simulator.make_move(action)
        if self.winner == 3:
            self.simulator.done = True
        if self.winner == self.current_player_num:
            r = 1
        elif self.winner == (self.current_player_num + 1) % 2:
            r = -1
        else:
            r = 0
        reward = [-r, -r]
        reward[self.current_player_num] = r
        self.current_player_num = (self.current_player_num + 1) % 2
        self.invert_board()
        return self.simulator.board, reward, self.simulator.done, {}

    def invert_board(self):
        """The reference transposes and sign-swaps simulator.board (HexSingleGame.py:265-271). The device keeps one board
        and derives the view; the game object only counts the flips (see DeviceGame.board)."""
        self.simulator._ref_flipped = not self.simulator._ref_flipped

    def render(self, mode="ansi", close=False):
        _print_board(self.simulator.board, 0, -1)


# ---------------------------------------------------------------------------------------------------- self-play wrapper
def selfplay_wrapper(env):
    """SelfplayWrapper.py:37-208: `env` (HexEnvB) plus an opponent that answers every agent ply."""

    class SelfPlayEnv(env):
        def __init__(self, base_model=BaseRandomPolicy(), scores=np.zeros(20), play_gui=False, board_size=5, buffer_size=20,
                     sample_board=False, prob_model=None, agent_player_num=None):
            if play_gui:
                raise NotImplementedError("play_gui needs the pygame GUI of the reference, which is out of scope")
            super(SelfPlayEnv, self).__init__(board_size=board_size, sample_board=sample_board)
            self.agent_player_num = agent_player_num
            self.calculate_probs = False
            if type(base_model) != BaseRandomPolicy:
                self.opponent_models = np.array([OpponentPolicy(base_model) for _ in range(buffer_size)])
                self.opponent_scores = scores
                base_model = OpponentPolicy(base_model)
            else:
                self.opponent_models = np.array([BaseRandomPolicy() for _ in range(buffer_size)])
                self.opponent_scores = np.zeros(buffer_size)
            self.best_model = base_model
            self.best_score = np.max(self.opponent_scores)
            self.best_mean_reward = -np.inf
            self.eval_state = False
            self.eval_episode = 0
            self.play_gui = False

        def reset(self, seed=None, options=None):
            super(SelfPlayEnv, self).reset()
            if self.agent_player_num is None:
                self.agent_player_num = random.randint(0, 1)
            self.setup_opponents()
            if self.current_player_num != self.agent_player_num:
                self.continue_game()
            info = {"state": self.simulator.board, "last_move_opponent": None, "last_move_player": None}
            return self.simulator.board, info

        def setup_opponents(self):
            if self.eval_state:
                if self.eval_episode <= len(self.opponent_models) - 1:
                    self.opponent_model = self.opponent_models[self.eval_episode]
                    self.eval_episode += 1
                return None
            rv = random.uniform(0, 1)
            if rv < 0.8:
                self.opponent_model = self.best_model
            else:
                self.opponent_model = self.opponent_models[int(random.random() * len(self.opponent_models))]

        def append_opponent_model(self, opponent_model, best_model=False, mean_reward=None):
            new_opponent = OpponentPolicy(opponent_model)
            if best_model:
                self.best_model = new_opponent
                self.best_mean_reward = mean_reward
            self.opponent_models.append(new_opponent)  # AttributeError on the ndarray pool, exactly like the reference

        def get_best_mean_reward(self):
            return self.best_mean_reward

        def set_eval(self, eval_state):
            self.eval_episode = 0
            self.eval_state = eval_state
            assert len(self.opponent_models) == len(self.opponent_scores)

        def get_scores(self):
            return self.opponent_scores

        def set_opponent_model(self, index, model, score):
            model = OpponentPolicy(model)
            self.opponent_models[index] = model
            self.opponent_scores[index] = score
            if score > self.best_score:
                self.best_model = model
                self.best_score = score

        def get_opponent_models(self):
            return self.opponent_models

        def save_best_model(self):
            self.best_model.save_model("models/best_model_" + str(self.best_score))

        def continue_game(self):
            random.uniform(0, 1)  # the reference draws and ignores it (SelfplayWrapper.py:159)
            action = self.opponent_model.choose_action(self.simulator.board, self.legal_actions())
            observation, reward, done, _ = super(SelfPlayEnv, self).step(action)
            return observation, reward, done, None

        def step(self, action):
            observation, reward, done, _ = super(SelfPlayEnv, self).step(action)
            if not done:
                observation, reward, done, _ = self.continue_game()
            return observation, reward[self.agent_player_num], done, False, {}

    return SelfPlayEnv
