"""Device-resident rollout feed for MaskablePPO-style training (BASELINE config 4).

The reference collects rollouts with stable-baselines3: `MaskablePPO.collect_rollouts` steps ONE Python env, asks it for its
action mask, samples from a masked categorical and appends to a `MaskableRolloutBuffer`
(scripts/experiments/6x6_MLP-default_lr-0.0003.py:34-47; SB3 2.2.1 / sb3-contrib from memory - not installed here, so the
buffer layout is duck-typed: obs [T,G,N,N], action_masks [T,G,C], actions, rewards, episode_starts, values, log_probs,
advantages, returns). Here the whole loop stays on the GPU: the fused step kernel writes observations and masks straight
into the buffer's slices, `masked_sample` (hexb_masked_sample, one warp per game) replaces the distribution object, and
nothing crosses PCIe until the learner wants a scalar.
"""
import ctypes

import torch

from . import _native
from ._native import check
from .batch import HexBatch


def masked_sample(logits, mask, u=None, generator=None, want_entropy=False, actions=None, logp=None):
    """Sample one legal action per row. logits f32[G,C], mask u8/bool[G,C] (1 = legal), u f64[G] uniforms (drawn with
    torch if None). Returns (actions i32[G], log_prob f32[G][, entropy f32[G]]); `actions` / `logp` may be given (e.g. slices
    of a rollout buffer) and are then written in place."""
    if not logits.is_cuda:
        raise RuntimeError("masked_sample runs on the GPU only (no CPU fallback)")
    G, C = logits.shape
    logits = logits.contiguous().float()
    mask = mask.contiguous()
    if mask.dtype == torch.bool:
        mask = mask.view(torch.uint8)
    if u is None:
        u = torch.rand(G, dtype=torch.float64, device=logits.device, generator=generator)
    u = u.contiguous().double()
    for name, t, dt in (("actions", actions, torch.int32), ("logp", logp, torch.float32)):
        if t is not None and (tuple(t.shape) != (G,) or t.dtype != dt or not t.is_contiguous() or t.device != logits.device):
            raise ValueError("%s must be a contiguous %s tensor of shape (%d,) on %s" % (name, dt, G, logits.device))
    actions = torch.empty(G, dtype=torch.int32, device=logits.device) if actions is None else actions
    logp = torch.empty(G, dtype=torch.float32, device=logits.device) if logp is None else logp
    ent = torch.empty(G, dtype=torch.float32, device=logits.device) if want_entropy else None
    p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    with torch.cuda.device(logits.device):
        check(_native.lib().hexb_masked_sample(p(logits), p(mask), p(u), G, C, p(actions), p(logp), p(ent), logits.device.index,
                                               ctypes.c_void_p(torch.cuda.current_stream(logits.device).cuda_stream)))
    return (actions, logp, ent) if want_entropy else (actions, logp)


def gae(rewards, values, dones, gamma=0.99, gae_lambda=0.95, advantages=None, returns=None):
    """GAE(lambda) of a [T,G] rollout in ONE kernel (hexb_gae, thread per game, backward over T) - what SB3's
    RolloutBuffer.compute_returns_and_advantage computes. rewards f32[T,G], values f32[T+1,G], dones u8[T,G] (the episode ended
    in step t). Returns (advantages f32[T,G], returns f32[T,G]), written into the given tensors when passed."""
    if not rewards.is_cuda:
        raise RuntimeError("gae runs on the GPU only (no CPU fallback)")
    T, G = rewards.shape
    dev = rewards.device
    for name, t, shape, dt in (("rewards", rewards, (T, G), torch.float32), ("values", values, (T + 1, G), torch.float32),
                               ("dones", dones, (T, G), torch.uint8)):
        if tuple(t.shape) != shape or t.dtype != dt or not t.is_contiguous() or t.device != dev:
            raise ValueError("%s must be a contiguous %s tensor of shape %s on %s" % (name, dt, shape, dev))
    advantages = torch.empty((T, G), dtype=torch.float32, device=dev) if advantages is None else advantages
    returns = torch.empty((T, G), dtype=torch.float32, device=dev) if returns is None else returns
    for name, t in (("advantages", advantages), ("returns", returns)):
        if tuple(t.shape) != (T, G) or t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev:
            raise ValueError("%s must be a contiguous float32 tensor of shape %s on %s" % (name, (T, G), dev))
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    with torch.cuda.device(dev):
        check(_native.lib().hexb_gae(p(rewards), p(values), p(dones), T, G, float(gamma), float(gae_lambda), p(advantages), p(returns),
                                     dev.index, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return advantages, returns


class RolloutBuffer(object):
    """[T, G] rollout storage on the device in MaskableRolloutBuffer's layout. Observations have the batch's obs_dtype: create
    the HexBatch with obs_dtype=torch.float32 and the step kernel writes the float32 observations the policy network reads
    (SB3's buffer is float32 too), with int8 they are the bytes the kernel computes in and the policy call converts."""

    def __init__(self, n_steps, batch: HexBatch, gamma=0.99, gae_lambda=0.95):
        T, G, N, C, dev = n_steps, batch.G, batch.N, batch.C, batch.device
        self.T, self.G, self.gamma, self.gae_lambda = T, G, gamma, gae_lambda
        self.obs = torch.zeros((T + 1, G, N, N), dtype=batch.obs_dtype, device=dev)  # slot T = the observation after the last step
        self.action_masks = torch.zeros((T + 1, G, C), dtype=torch.uint8, device=dev)
        self.actions = torch.zeros((T, G), dtype=torch.int32, device=dev)
        self.rewards = torch.zeros((T, G), dtype=torch.float32, device=dev)
        self.dones = torch.zeros((T, G), dtype=torch.uint8, device=dev)
        self.episode_starts = torch.zeros((T + 1, G), dtype=torch.float32, device=dev)
        self.values = torch.zeros((T + 1, G), dtype=torch.float32, device=dev)
        self.log_probs = torch.zeros((T, G), dtype=torch.float32, device=dev)
        self.advantages = torch.zeros((T, G), dtype=torch.float32, device=dev)
        self.returns = torch.zeros((T, G), dtype=torch.float32, device=dev)

    def compute_returns_and_advantage(self):
        """GAE(lambda) exactly as SB3's RolloutBuffer: a finished episode (auto-reset) cuts the bootstrap. One hexb_gae launch."""
        gae(self.rewards, self.values, self.dones, self.gamma, self.gae_lambda, self.advantages, self.returns)

    def compute_returns_and_advantage_torch(self):
        """The same recurrence as T eager PyTorch steps: the float32 reference hexb_gae is tested against (tests/test_gpu_rollout.py)."""
        last = torch.zeros(self.G, dtype=torch.float32, device=self.values.device)
        for t in reversed(range(self.T)):
            nonterminal = 1.0 - self.episode_starts[t + 1]
            delta = self.rewards[t] + self.gamma * self.values[t + 1] * nonterminal - self.values[t]
            last = delta + self.gamma * self.gae_lambda * nonterminal * last
            self.advantages[t] = last
        torch.add(self.advantages, self.values[:-1], out=self.returns)

    def minibatches(self, batch_size, generator=None):
        n = self.T * self.G
        perm = torch.randperm(n, device=self.obs.device, generator=generator)
        flat = dict(obs=self.obs[:-1].reshape(n, *self.obs.shape[2:]), action_masks=self.action_masks[:-1].reshape(n, -1),
                    actions=self.actions.reshape(n), values=self.values[:-1].reshape(n), log_probs=self.log_probs.reshape(n),
                    advantages=self.advantages.reshape(n), returns=self.returns.reshape(n))
        for s in range(0, n, batch_size):
            idx = perm[s:s + batch_size]
            yield {k: v[idx] for k, v in flat.items()}


class RolloutCollector(object):
    """collect(): T fused env steps of all games with actions sampled from `policy` under the legal-action mask.
    `policy(obs_f32[G,N,N]) -> (logits f32[G,C], values f32[G])` is any torch callable living on the same GPU."""

    def __init__(self, batch: HexBatch, n_steps, gamma=0.99, gae_lambda=0.95, seed=0, extra_generators=()):
        """extra_generators: torch.Generator objects that policy / opponent_fn draw from (needed for use_graph=True: a CUDA
        graph can only advance generators registered with it)."""
        self.batch, self.buf = batch, RolloutBuffer(n_steps, batch, gamma, gae_lambda)
        self.extra_generators = list(extra_generators)
        self.gen = torch.Generator(device=batch.device)
        self.gen.manual_seed(seed)
        self._started = False
        self._graph, self._graph_key = None, None

    def restart(self):
        """The games were reset or stepped behind the collector's back (an evaluation pass, opponents.evaluate_pool): the next
        collect() starts from a fresh reset instead of carrying the last observation over. A captured graph stays valid."""
        self._started = False

    def collect(self, policy, opponent_fn=None, use_graph=False):
        """opponent_fn: only for manual_opponent batches - the learned opponent (see HexBatch.step_with_opponent).

        use_graph=True: the whole rollout (T x [policy forward, masked sampling, env step] + GAE) is captured into ONE CUDA
        graph the second time it is collected and replayed from then on - at 4,096 games a step is a handful of 5-10 us
        kernels, so issuing them from Python costs several times their run time. The graph reads the policy's parameters in
        place (optimizer steps and load_state_dict update them in place), draws from this collector's generator, and
        writes the same buffer tensors every time. `policy` / `opponent_fn` must be the same objects on every call."""
        b, buf = self.batch, self.buf
        if b.manual_opponent and opponent_fn is None:
            raise ValueError("a manual_opponent batch needs opponent_fn")
        if not self._started:
            b.reset(obs=buf.obs[0], mask=buf.action_masks[0])
            if b.manual_opponent:
                b.opponent_opening(opponent_fn)
                b.encode(0, obs=buf.obs[0], mask=buf.action_masks[0])
            buf.episode_starts[0] = 1.0
            self._started = True
            self._rollout(policy, opponent_fn, carry=False)      # the first rollout always runs eagerly (it also warms cuBLAS up)
            return buf
        if not use_graph:
            self._rollout(policy, opponent_fn, carry=True)
            return buf
        key = (id(policy), id(opponent_fn), getattr(opponent_fn, "version", 0))   # an OpponentPool counts its changes
        if self._graph is None or self._graph_key != key:
            g = torch.cuda.CUDAGraph()
            for gen in [self.gen] + self.extra_generators:
                g.register_generator_state(gen)
            torch.cuda.synchronize(b.device)
            with torch.cuda.graph(g):
                self._rollout(policy, opponent_fn, carry=True)
            self._graph, self._graph_key = g, key
        self._graph.replay()
        return buf

    def _rollout(self, policy, opponent_fn, carry):
        b, buf = self.batch, self.buf
        if carry:  # continue from where the previous rollout stopped
            buf.obs[0].copy_(buf.obs[-1])
            buf.action_masks[0].copy_(buf.action_masks[-1])
            buf.episode_starts[0].copy_(buf.episode_starts[-1])
        with torch.no_grad():
            for t in range(buf.T):
                obs = buf.obs[t]
                logits, values = policy(obs if obs.dtype == torch.float32 else obs.float())
                actions, _ = masked_sample(logits, buf.action_masks[t], generator=self.gen, actions=buf.actions[t], logp=buf.log_probs[t])
                buf.values[t].copy_(values)
                if b.manual_opponent:
                    o = b.step_with_opponent(actions, opponent_fn, obs=buf.obs[t + 1], mask=buf.action_masks[t + 1])
                    buf.rewards[t], buf.dones[t] = o["reward"], o["done"]
                else:
                    b.step(actions, obs=buf.obs[t + 1], mask=buf.action_masks[t + 1], reward=buf.rewards[t], done=buf.dones[t])
            obs = buf.obs[buf.T]
            buf.values[buf.T].copy_(policy(obs if obs.dtype == torch.float32 else obs.float())[1])
            buf.episode_starts[1:].copy_(buf.dones)     # SB3's layout: a step starts an episode iff the previous one ended it
        buf.compute_returns_and_advantage()
