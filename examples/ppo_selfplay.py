#!/usr/bin/env python
"""BASELINE config 4: 6x6 MaskablePPO-style self-play training with the batched env feeding a device-resident rollout buffer.

What scripts/experiments/6x6_MLP-default_lr-0.0003.py of the reference does with SB3 (MaskablePPO("MlpPolicy"), lr 3e-4,
gamma 0.99, lambda 0.95, clip 0.2, 10 epochs, default MlpPolicy = separate 2x64 tanh nets for pi and vf - shapes read from
models/6x6_MLP-default_lr-0.0003_71), here with G games collected in lockstep on one GPU. The opponent is the random policy
(BaseRandomPolicy); SB3 itself is not installed, so the learner below is a plain torch restatement of the PPO update.

    python examples/ppo_selfplay.py --board 6 --games 4096 --n-steps 128 --iters 20
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

from hex_gym_env_b200 import AGENT_RANDOM, VARIANT_B, HexBatch  # noqa: E402
from hex_gym_env_b200.rollout import RolloutCollector, masked_sample  # noqa: E402


class MlpPolicy(nn.Module):
    def __init__(self, cells, hidden=64):
        super().__init__()
        self.pi = nn.Sequential(nn.Flatten(), nn.Linear(cells, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh(), nn.Linear(hidden, cells))
        self.vf = nn.Sequential(nn.Flatten(), nn.Linear(cells, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh(), nn.Linear(hidden, 1))

    def forward(self, obs):
        return self.pi(obs), self.vf(obs).squeeze(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--board", type=int, default=6)
    ap.add_argument("--games", type=int, default=4096)
    ap.add_argument("--n-steps", type=int, default=128)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--epochs", type=int, default=10)
    ap.add_argument("--minibatch", type=int, default=16384)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--opponent", choices=["random", "self", "pool"], default="random",
                    help="random: BaseRandomPolicy inside the fused step kernel; self: a frozen copy of the policy, refreshed every "
                         "--refresh iterations, played through the split step (hexb_half_step); pool: the reference's opponent buffer "
                         "(SelfplayWrapper.py:39-67,91-104 + SelfPlayCallback, EvaluationCallback.py:31-50): --pool-size frozen "
                         "snapshots, per-episode choice on the device (80 %% the best one), every --eval-every iterations an "
                         "evaluation against every entry and, if the learner scores, the worst entry replaced by it")
    ap.add_argument("--refresh", type=int, default=4)
    ap.add_argument("--pool-size", type=int, default=8)
    ap.add_argument("--eval-every", type=int, default=4)
    ap.add_argument("--obs-dtype", choices=["f32", "i8"], default="f32",
                    help="f32: the step kernel writes float32 observations straight into the rollout buffer (hexb_config.obs_dtype); "
                         "i8: int8 observations, converted for the network at every use")
    ap.add_argument("--graph", action="store_true", help="replay the whole rollout (policy forward + sampling + env step + GAE) as one CUDA graph")
    args = ap.parse_args()
    torch.manual_seed(args.seed)
    dev = torch.device("cuda", 0)
    env = HexBatch(args.board, args.games, variant=VARIANT_B, device=0, seed=args.seed, agent_mode=AGENT_RANDOM, auto_reset=True,
                   manual_opponent=(args.opponent != "random"), pool_size=args.pool_size if args.opponent == "pool" else 0,
                   obs_dtype=torch.float32 if args.obs_dtype == "f32" else torch.int8)
    policy = MlpPolicy(env.C).to(dev)
    frozen = MlpPolicy(env.C).to(dev)
    frozen.load_state_dict(policy.state_dict())
    ogen = torch.Generator(device=dev)
    ogen.manual_seed(args.seed + 99)

    def opponent_fn(obs, mask, to_move, opp_index):      # what OpponentPolicy.choose_action does, for all waiting games at once
        with torch.no_grad():
            logits, _ = frozen(obs.float())
            return masked_sample(logits, mask, generator=ogen)[0]

    pool = None
    if args.opponent == "pool":
        from hex_gym_env_b200.opponents import OpponentPool, StackedMlpOpponents, evaluate_pool

        # every pool entry (and the best model) is a slot of ONE stack of weights: a batched matmul per layer evaluates all of them
        # for all games; replacing an entry loads the learner's weights INTO its slot, so a captured CUDA graph stays valid
        def pi_linears():
            return [m for m in policy.pi if isinstance(m, nn.Linear)]

        stack = StackedMlpOpponents((env.C, 64, 64, env.C), args.pool_size + 1, device=dev, generator=ogen)
        for sl in range(args.pool_size + 1):
            stack.load(sl, pi_linears())
        pool = OpponentPool(stack.entry(0), buffer_size=args.pool_size, batch=env)
        for k in range(args.pool_size):
            pool.set_opponent_model(k, stack.entry(k + 1), 0.0)
        opponent_fn = pool
        egen = torch.Generator(device=dev)
        egen.manual_seed(args.seed + 5)

        def greedy_agent(obs, mask):                       # the evaluation plays the learner's sampled policy, like training
            with torch.no_grad():
                return masked_sample(policy(obs.float())[0], mask, generator=egen)[0]

    opt = torch.optim.Adam(policy.parameters(), lr=args.lr, eps=1e-5, capturable=args.graph)
    col = RolloutCollector(env, args.n_steps, gamma=0.99, gae_lambda=0.95, seed=args.seed, extra_generators=[ogen])
    buf = col.buf
    n = buf.T * buf.G
    flat = dict(obs=buf.obs[:-1].reshape(n, *buf.obs.shape[2:]), action_masks=buf.action_masks[:-1].reshape(n, -1),
                actions=buf.actions.reshape(n), log_probs=buf.log_probs.reshape(n), advantages=buf.advantages.reshape(n),
                returns=buf.returns.reshape(n))                              # views of the collector's (static) tensors
    idx = torch.zeros(min(args.minibatch, n), dtype=torch.long, device=dev)  # the minibatch's sample indices
    loss_out = torch.zeros((), device=dev)

    def update_minibatch():
        """One clipped-PPO optimizer step on the samples idx points at (sb3_contrib MaskablePPO.train, one minibatch)."""
        mb = {k: v[idx] for k, v in flat.items()}
        logits, values = policy(mb["obs"].float())
        logits = logits.masked_fill(mb["action_masks"] == 0, -1e8)          # sb3_contrib's HUGE_NEG masking
        logp_all = torch.log_softmax(logits, dim=-1)
        logp = logp_all.gather(1, mb["actions"].long().unsqueeze(1)).squeeze(1)
        adv = mb["advantages"]
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        ratio = torch.exp(logp - mb["log_probs"])
        pg = -torch.min(adv * ratio, adv * torch.clamp(ratio, 0.8, 1.2)).mean()
        vloss = torch.nn.functional.mse_loss(values, mb["returns"])
        ent = -(torch.exp(logp_all) * logp_all.masked_fill(mb["action_masks"] == 0, 0.0)).sum(-1).mean()
        loss = pg + 0.5 * vloss - 0.0 * ent
        loss.backward()
        nn.utils.clip_grad_norm_(policy.parameters(), 0.5)
        opt.step()
        loss_out.copy_(loss.detach())

    update_graph = None
    pgen = torch.Generator(device=dev)
    pgen.manual_seed(args.seed + 7)
    prev = env.stats().cpu()
    for it in range(args.iters):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if args.opponent == "self" and it and it % args.refresh == 0:
            frozen.load_state_dict(policy.state_dict())
        if pool is not None and it and it % args.eval_every == 0:
            ev = evaluate_pool(env, pool, greedy_agent)      # set_eval(True): every game meets every entry once; set_eval(False)
            def into_slot(i):                                # the learner replaces the worst entry: its weights go into that slot
                stack.load(i + 1, pi_linears())
                return stack.entry(i + 1)

            score, idx_rep = pool.consider(policy, ev["mean_reward"], place=into_slot)
            print(json.dumps({"iter": it, "eval_mean_reward": ev["mean_reward"], "eval_episodes": ev["episodes"], "score": score,
                              "replaced_entry": idx_rep, "pool_scores": [round(float(x), 4) for x in pool.get_scores()]}))
            col.restart()                                    # the evaluation reset the games: the next rollout starts from a reset
        col.collect(policy, opponent_fn if args.opponent != "random" else None, use_graph=args.graph)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if args.graph and it == 1 and n % idx.numel() == 0:
            # the first iteration's 320 eager optimizer steps were the warm-up; from now on one minibatch step = one graph replay
            update_graph = torch.cuda.CUDAGraph()
            opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(update_graph):
                update_minibatch()
        for _ in range(args.epochs):
            perm = torch.randperm(n, device=dev, generator=pgen)
            for s0 in range(0, n, idx.numel()):
                if update_graph is not None:
                    idx.copy_(perm[s0:s0 + idx.numel()])
                    update_graph.replay()
                else:
                    cur = perm[s0:s0 + idx.numel()]
                    if cur.numel() != idx.numel():
                        continue                                              # (ragged tail: dropped, like drop_last)
                    idx.copy_(cur)
                    opt.zero_grad(set_to_none=True)
                    update_minibatch()
        loss = loss_out
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        st = env.stats().cpu()
        d = (st - prev).tolist()
        prev = st
        steps = args.games * args.n_steps
        print(json.dumps({"iter": it, "collect_env_steps_per_sec": steps / (t1 - t0), "end_to_end_env_steps_per_sec": steps / (t2 - t0),
                          "episodes": d[0], "agent_win_rate": d[3] / max(d[0], 1), "invalid_rate": d[5] / max(d[0], 1),
                          "mean_episode_plies": d[4] / max(d[0], 1), "loss": float(loss)}))


if __name__ == "__main__":
    main()
