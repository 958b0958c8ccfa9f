"""TEST INFRASTRUCTURE ONLY - a deterministic stand-in for a stable-baselines opponent model.

`ScriptedModel.predict(board, deterministic, action_masks)` has the signature OpponentPolicy calls
(minihex/SelfplayWrapper.py:30-32). Its choice is a pure function of the board the env shows it (the side-to-move view)
and the legal-action mask, so it pins both the caller-driven-opponent path and the opponent's observation."""
import numpy as np


def scripted_choice(board, mask):
    flat = np.asarray(board).reshape(-1)
    legal = np.flatnonzero(np.asarray(mask).reshape(-1))
    own = np.flatnonzero(flat == -1)
    other = np.flatnonzero(flat == 1)
    h = (int(own.sum()) * 31 + int(other.sum()) * 17 + len(own) * 7 + 3) % len(legal)
    return int(legal[h])


class ScriptedModel(object):
    def __init__(self, ident, log):
        self.ident, self.log = ident, log

    def predict(self, board, deterministic=False, action_masks=None):
        a = scripted_choice(board, action_masks)
        self.log.append((self.ident, a))
        return a, None

    def save(self, path):
        return None


def scripted_choice_a(board, mask):
    """The same rule for a variant-A board (codes BLACK 0 / WHITE 1 / EMPTY 2) as HexEnv.opponent_predict shows it to its model
    (minihex/HexGame.py:354-359: the transposed, colour-swapped board and the mask of that view)."""
    flat = np.asarray(board).reshape(-1)
    legal = np.flatnonzero(np.asarray(mask).reshape(-1))
    own = np.flatnonzero(flat == 0)
    other = np.flatnonzero(flat == 1)
    h = (int(own.sum()) * 31 + int(other.sum()) * 17 + len(own) * 7 + 3) % len(legal)
    return int(legal[h])


class ScriptedModelA(object):
    """predict(state, deterministic=True, action_masks=...) as HexEnv.opponent_predict calls it; logs (action, mask) per call."""

    def __init__(self, log):
        self.log = log

    def predict(self, board, deterministic=False, action_masks=None):
        a = scripted_choice_a(board, action_masks)
        self.log.append((a, np.asarray(action_masks).astype(np.uint8).copy(), np.asarray(board).astype(np.int8).copy()))
        return a, None
