"""The opponent pool of SelfPlayEnv for a whole batch of games.

In the reference every env object owns a buffer of opponent models (minihex/SelfplayWrapper.py:39-67), picks one per episode in
setup_opponents (:91-104: 80 % the best model, else a uniformly drawn entry; while evaluating, episode k meets entry k) and asks it
for every reply through OpponentPolicy.choose_action (:26-35, called from continue_game :161-163). SelfPlayCallback
(minihex/EvaluationCallback.py:31-50) switches evaluation on and off and replaces the worst entry by the learner when it scores.

Here the per-episode choice is made on the device at every restart (hexb_core.cuh reset_game -> HexBatch.opp_index, int32[G]:
-1 = best model, k = entry k; hexb_set_eval for the evaluation cycle), bit-exact against the unmodified reference
(tests/golden/oppmodel_*.npz, evalpool_*.npz). This module is the host side of it:

  * OpponentPool keeps the reference's bookkeeping under the reference's names (opponent_models, opponent_scores, best_model,
    best_score, best_mean_reward, eval_state; append_opponent_model, set_opponent_model, get_scores, get_opponent_models,
    set_eval, get_best_mean_reward, save_best_model) and
  * is itself the `opponent_fn(obs, mask, to_move, opp_index) -> actions` that HexBatch.step_with_opponent / opponent_opening /
    RolloutCollector.collect take: it groups the waiting games by their pool entry and asks every DISTINCT model once for its
    group - OpponentPolicy.choose_action for all games at once.

A pool entry is a batched policy: a callable `model(obs, mask) -> int32[n]` of actions in the mover's own view for n games
(obs [n,N,N] in the batch's obs_dtype, mask u8[n,C]; the batched form of `model.predict(board, action_masks=mask)[0]`).
No game logic lives here; without a HexBatch the class only keeps books and dispatches (that part is tested on the CPU).
"""
import math
import random as _random

import numpy as np
import torch

__all__ = ["OpponentPool", "StackedMlpOpponents", "evaluate_pool"]


class OpponentPool(object):
    def __init__(self, base_model, buffer_size=20, scores=None, batch=None, dense=True):
        """base_model fills all buffer_size entries and is the best model, as SelfPlayEnv.__init__ does (:56-63).
        batch: the manual_opponent HexBatch whose games meet this pool (its pool_size must equal buffer_size), or None.
        dense=True asks each distinct model for ALL games and merges by opp_index (no host synchronisation: usable inside a CUDA
        graph); dense=False gathers each model's games first (one device->host synchronisation per call, less network work)."""
        if buffer_size < 1:
            raise ValueError("buffer_size must be at least 1")
        self.opponent_models = [base_model for _ in range(buffer_size)]
        self.opponent_scores = np.zeros(buffer_size) if scores is None else np.asarray(scores, dtype=np.float64).copy()
        if len(self.opponent_scores) != buffer_size:
            raise ValueError("scores must have buffer_size entries")
        self.best_model = base_model
        self.best_score = float(np.max(self.opponent_scores))
        self.best_mean_reward = -np.inf
        self.eval_state = False
        self.dense = bool(dense)
        self.version = 0     # bumped whenever an entry or the best model changes: a CUDA graph that captured __call__ is stale then
        self.batch = None
        if batch is not None:
            self.bind(batch)

    # ------------------------------------------------------------------ the batch this pool plays in
    def bind(self, batch):
        if not getattr(batch, "manual_opponent", False):
            raise ValueError("an opponent pool needs a HexBatch created with manual_opponent=True")
        if int(batch.cfg.pool_size) != len(self.opponent_models):
            raise ValueError("HexBatch pool_size (%d) and the pool's length (%d) differ: the device draws the entry index"
                             % (int(batch.cfg.pool_size), len(self.opponent_models)))
        self.batch = batch
        if bool(batch.cfg.eval_state) != self.eval_state:
            batch.set_eval(self.eval_state)
        return self

    # ------------------------------------------------------------------ the reference's pool API (SelfplayWrapper.py:106-144)
    def append_opponent_model(self, opponent_model, best_model=False, mean_reward=None):
        """:106-112. The device draws indices below the pool size its HexBatch was created with, so a bound pool cannot grow."""
        if self.batch is not None:
            raise ValueError("the pool of a bound batch has a fixed length (hexb_config.pool_size); use set_opponent_model")
        if best_model:
            self.best_model = opponent_model
            self.best_mean_reward = mean_reward
        self.opponent_models.append(opponent_model)
        self._changed()

    def get_best_mean_reward(self):
        return self.best_mean_reward

    def set_eval(self, eval_state):
        """:117-120; on the device hexb_set_eval: restarts draw nothing and walk the pool, episode k of a game meets entry k."""
        assert len(self.opponent_models) == len(self.opponent_scores)
        self.eval_state = bool(eval_state)
        if self.batch is not None:
            self.batch.set_eval(self.eval_state)

    def get_scores(self):
        return self.opponent_scores

    def set_opponent_model(self, index, model, score):
        """:125-136."""
        self.opponent_models[index] = model
        self.opponent_scores[index] = score
        if score > self.best_score:
            self.best_model = model
            self.best_score = score
        self._changed()

    def get_opponent_models(self):
        return self.opponent_models

    def save_best_model(self, path=None):
        """:141-143: models/best_model_<score>; entries without a save method (BaseRandomPolicy.save_model returns None) are skipped."""
        name = path or ("models/best_model_" + str(self.best_score))
        for attr in ("save_model", "save"):
            fn = getattr(self.best_model, attr, None)
            if callable(fn):
                return fn(name)
        return None

    def consider(self, model, last_mean_reward, rng=_random, place=None):
        """SelfPlayCallback._on_step after an evaluation (EvaluationCallback.py:35-50): score = mean reward * exp(mean(scores) - 1);
        a learner that won on average (mean reward > 0) and beats the worst score replaces one of the worst entries, chosen with
        random.choice. Returns (score, replaced index or None). place(index) -> the object to store in the chosen slot instead of
        `model` (e.g. the slot's own network after the learner's weights were loaded into it, which keeps a captured CUDA graph valid)."""
        scores = self.opponent_scores
        score = float(last_mean_reward * math.exp(float(np.mean(scores)) - 1.0))
        replaced = None
        if last_mean_reward > 0 and score > np.min(scores):
            worst = np.flatnonzero(scores == np.min(scores))
            replaced = int(rng.choice(list(worst)))
            self.set_opponent_model(replaced, place(replaced) if place is not None else model, score)
        return score, replaced

    # ------------------------------------------------------------------ OpponentPolicy.choose_action for every waiting game
    def groups(self):
        """[(model, [entry indices])] with one item per DISTINCT model object, -1 standing for best_model."""
        out = []
        for k, m in [(-1, self.best_model)] + list(enumerate(self.opponent_models)):
            for item in out:
                if item[0] is m:
                    item[1].append(k)
                    break
            else:
                out.append((m, [k]))
        return out

    def _stack_slots(self, device):
        """When every entry (and the best model) is a slot of one StackedMlpOpponents: (stack, long[pool + 1] slot of index k + 1)."""
        models = [self.best_model] + list(self.opponent_models)
        stack = getattr(models[0], "stack", None)
        if stack is None or any(getattr(m, "stack", None) is not stack for m in models):
            return None, None
        cache = self.__dict__.setdefault("_slots", {})
        hit = cache.get(str(device))
        if hit is None or hit[0] != self.version:
            hit = (self.version, torch.tensor([m.slot for m in models], dtype=torch.long, device=device))
            cache[str(device)] = hit
        return stack, hit[1]

    def _changed(self):
        """An entry or the best model changed: rebuild the slot table right away for the devices it is used on (a host-to-device
        copy, which must not wait for the next __call__ - that one may run inside a CUDA-graph capture)."""
        self.version += 1
        for dev in list(self.__dict__.get("_slots", {})):
            self._stack_slots(torch.device(dev))

    def __call__(self, obs, mask, to_move, opp_index):
        stack, slot_of = self._stack_slots(opp_index.device)
        if stack is not None:   # one batched evaluation of all entries, each game keeps its own (games not waiting are ignored anyway)
            return stack.actions(obs, mask, slot_of[opp_index.long() + 1])
        actions = torch.zeros(opp_index.shape[0], dtype=torch.int32, device=opp_index.device)
        waiting = to_move == 1
        for model, entries in self.groups():
            sel = opp_index == entries[0]
            for k in entries[1:]:
                sel = sel | (opp_index == k)
            sel = sel & waiting
            if self.dense:
                a = model(obs, mask)
                actions = torch.where(sel, a.to(torch.int32), actions)
            else:
                idx = torch.nonzero(sel).squeeze(1)
                if idx.numel():
                    actions[idx] = model(obs[idx], mask[idx]).to(torch.int32)
        return actions


class StackedMlpOpponents(object):
    """Every pool entry as ONE set of stacked weights: S policy networks of the same architecture (the reference's pool holds
    MaskablePPO MlpPolicy snapshots, all `pi` = Linear-tanh-Linear-tanh-Linear, scripts/experiments/*.py:40) live in tensors
    W_l[S, in_l, out_l], b_l[S, out_l], so one batched matmul per layer evaluates all of them for all games and a gather keeps
    each game's own entry - a handful of launches per opponent pass whatever the pool size, instead of one forward per entry.
    Replacing an entry copies the learner's weights INTO its slot, so a captured CUDA graph keeps reading the right memory.

    entry(slot) is that slot as a pool entry (a batched policy on its own, and recognised by OpponentPool, which then asks the
    stack once for all games). Actions are sampled with hexb_masked_sample (OpponentPolicy.choose_action calls
    predict(deterministic=False)); deterministic=True takes the legal arg-max instead (pure torch, also what the CPU tests use)."""

    class Entry(object):
        def __init__(self, stack, slot):
            self.stack, self.slot = stack, int(slot)

        def __call__(self, obs, mask):
            slots = torch.full((obs.shape[0],), self.slot, dtype=torch.long, device=obs.device)
            return self.stack.actions(obs, mask, slots)

    def __init__(self, layer_sizes, slots, device=None, deterministic=False, generator=None):
        self.sizes, self.S = [int(x) for x in layer_sizes], int(slots)
        self.deterministic, self.generator = bool(deterministic), generator
        self.W = [torch.zeros(self.S, i, o, device=device) for i, o in zip(self.sizes[:-1], self.sizes[1:])]
        self.b = [torch.zeros(self.S, o, device=device) for o in self.sizes[1:]]

    def entry(self, slot):
        if not 0 <= slot < self.S:
            raise IndexError("slot %r of %d" % (slot, self.S))
        return StackedMlpOpponents.Entry(self, slot)

    def load(self, slot, linears):
        """Copy a network into a slot, in place. linears: its torch.nn.Linear layers in order (e.g. [m for m in policy.pi if
        isinstance(m, nn.Linear)]), or (weight[out,in], bias[out]) pairs."""
        linears = list(linears)
        if len(linears) != len(self.W):
            raise ValueError("expected %d linear layers, got %d" % (len(self.W), len(linears)))
        with torch.no_grad():
            for l, lin in enumerate(linears):
                w, b = (lin.weight, lin.bias) if hasattr(lin, "weight") else lin
                if tuple(w.shape) != (self.sizes[l + 1], self.sizes[l]):
                    raise ValueError("layer %d has shape %s, the stack holds %s" % (l, tuple(w.shape), (self.sizes[l + 1], self.sizes[l])))
                self.W[l][slot].copy_(w.detach().t())
                self.b[l][slot].copy_(b.detach())

    def logits(self, obs, slots):
        """f32[G, out]: the network of slot slots[g] applied to game g; every slot is evaluated for every game (dense)."""
        G = obs.shape[0]
        h = obs.reshape(G, -1).to(self.W[0].dtype).unsqueeze(0).expand(self.S, G, self.sizes[0])
        for l in range(len(self.W)):
            h = torch.baddbmm(self.b[l].unsqueeze(1), h, self.W[l])
            if l + 1 < len(self.W):
                h = torch.tanh(h)
        return h[slots.long(), torch.arange(G, device=h.device)]

    def actions(self, obs, mask, slots):
        with torch.no_grad():
            lg = self.logits(obs, slots)
            if self.deterministic:
                return lg.masked_fill(mask == 0, float("-inf")).argmax(dim=1).to(torch.int32)
            from .rollout import masked_sample
            return masked_sample(lg, mask, generator=self.generator)[0]


def evaluate_pool(batch, pool, agent_fn, max_steps=None):
    """The evaluation SelfPlayCallback runs (EvaluationCallback.py:31-33: set_eval(True); n_eval_episodes = buffer_size episodes,
    episode k against pool entry k; set_eval(False)), for every game of the batch at once: each game plays one episode against
    every entry. agent_fn(obs, mask) -> int32[G] actions. Returns dict(mean_reward = what evaluate_policy hands the callback as
    last_mean_reward, per_entry = float64[pool] mean reward against each entry, episodes, per_entry_episodes); the games are left
    freshly reset in training mode.
    One device->host read per step decides when every game has met every entry (max_steps bounds the loop)."""
    K = len(pool.opponent_models)
    G, dev = batch.G, batch.device
    if pool.batch is not batch:
        pool.bind(batch)
    pool.set_eval(True)
    obs, mask = batch.reset()
    batch.opponent_opening(pool)
    obs, mask = batch.encode(0)
    total = torch.zeros(K, dtype=torch.float64, device=dev)
    count = torch.zeros(K, dtype=torch.int64, device=dev)
    played = torch.zeros(G, dtype=torch.int64, device=dev)       # finished evaluation episodes per game
    limit = max_steps if max_steps is not None else 4 * K * batch.C
    for _ in range(limit):
        entry = batch.opp_index.clone().long()                   # the entry the running episode is played against
        out = batch.step_with_opponent(agent_fn(obs, mask), pool)
        obs, mask = out["obs"], out["mask"]
        fin = (out["done"] != 0) & (played < K) & (entry >= 0)
        total.index_add_(0, entry.clamp(min=0), torch.where(fin, out["reward"].double(), torch.zeros((), dtype=torch.float64, device=dev)))
        count.index_add_(0, entry.clamp(min=0), fin.long())
        played += (out["done"] != 0).long()
        if int((played >= K).all()):
            break
    pool.set_eval(False)
    batch.reset()
    batch.opponent_opening(pool)
    n = int(count.sum())
    per_entry = (total / count.clamp(min=1).double()).cpu().numpy()
    return dict(mean_reward=float(total.sum()) / max(n, 1), per_entry=per_entry, episodes=n, per_entry_episodes=count.cpu().numpy())
