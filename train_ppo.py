#!/usr/bin/env python
"""End-to-end PantheonRL-style PPO training off the GPU env.

    python train_ppo.py --json-path cfg.json            # same JSON schema as the reference trainer
    python train_ppo.py                                 # open-divider_tomato, 65 536 envs: delivers 100 % after ~13 s
    torchrun --nproc-per-node 8 train_ppo.py            # data parallel: --envs per GPU, gradients all-reduced (NCCL)

Ego PPO + partner PPO (the partner records and trains inside `env.step`, like PantheonRL's
OnPolicyAgent), both on the device; the Overcooked env is `OvercookedVecEnv` (one CUDA launch per
step, no host sync inside a rollout).  Prints one JSON line per log interval.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

from gym_comm_b200 import OvercookedVecEnv, create_arglist, namespace_from_dict
from gym_comm_b200.pantheon import BatchedOnPolicyAgent, GraphedRollout, PantheonVecEnv, collect_and_train
from gym_comm_b200.ppo import PPO, PPOConfig, RecurrentPPO, save_learner
from gym_comm_b200.sharding import rank_world, shard_seed


def main(argv=None, env_factory=None, learners_out=None):
    """`env_factory(ns, args)` lets the CPU tests put the emulated env under the same training loop;
    `learners_out` (a list) receives the ego and the partner learner."""
    ap = argparse.ArgumentParser()
    ap.add_argument("--json-path", default=None, help="env config JSON (reference trainer.py --json-path)")
    ap.add_argument("--level", default="open-divider_tomato")
    ap.add_argument("--max-num-timesteps", type=int, default=200)
    ap.add_argument("--num-communication", type=int, default=10)
    ap.add_argument("--envs", type=int, default=65536, help="envs per process (per GPU under torchrun)")
    ap.add_argument("--n-steps", type=int, default=32)
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--total-timesteps", type=int, default=0,
                    help="train for this many ego env-steps summed over all envs (SB3's `learn(total_timesteps)`, "
                         "trainer.py:121) instead of --iters; the JSON's own total_timesteps is not applied automatically")
    ap.add_argument("--batch-size", type=int, default=262144,
                    help="minibatch of the PPO update; with 65,536 envs x 32 steps a rollout holds 2 M samples, and large "
                         "minibatches keep the (memory-bound) update kernels busy: 262,144 x 2 epochs trains at 30 M "
                         "agent-steps/s and solves open-divider_tomato in 13 s, 65,536 x 4 epochs at 10 M / 21 s "
                         "(profiles/r2_train_n1.jsonl)")
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--clip-range", type=float, default=0.2)
    ap.add_argument("--ent-coef", type=float, default=0.01)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--reward-scale", type=float, default=0.1, help="learner-side reward scaling (1.0 = the reference's raw reward)")
    ap.add_argument("--recurrent", action="store_true",
                    help="LSTM actor-critic for both learners (the reference's RecurrentPPO, trainer.py:92-121)")
    ap.add_argument("--lstm-hidden", type=int, default=256)
    ap.add_argument("--no-graph", action="store_true",
                    help="issue every rollout step and every minibatch update from Python instead of replaying CUDA graphs")
    ap.add_argument("--log-every", type=int, default=10)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--eval-steps", type=int, default=0, help="after training: deterministic evaluation for this many steps")
    ap.add_argument("--save-dir", default=None,
                    help="directory for the two policies, `ppo_ego.pt` and `ppo_partner1.pt` (trainer.py:129-133); "
                         "evaluate_policy.py --ego-load / --alt-load reads them")
    args = ap.parse_args(argv)

    if args.json_path:
        ns = create_arglist(args.json_path)
    else:
        ns = namespace_from_dict(dict(level=args.level, num_agents=2, max_num_timesteps=args.max_num_timesteps,
                                      communication_on=True, num_communication=args.num_communication))
    # one process per GPU under torchrun: every rank owns `--envs` envs (its own shard seed) and its own rollout
    # buffers; the learners start from the same weights and average their gradients (ppo.average_gradients)
    rank, world, local_rank = rank_world()
    if world > 1:
        if args.device.startswith("cuda"):
            args.device = "cuda:%d" % local_rank
            torch.cuda.set_device(local_rank)
        if not dist.is_initialized():
            dist.init_process_group("nccl" if args.device.startswith("cuda") else "gloo")
    env_seed = shard_seed(args.seed, rank) if world > 1 else args.seed
    torch.manual_seed(args.seed + rank)                   # action sampling differs per shard
    args.env_seed = env_seed
    env = env_factory(ns, args) if env_factory is not None else \
        OvercookedVecEnv(ns, num_envs=args.envs, device=args.device, seed=env_seed, auto_reset=True)
    h = getattr(ns, "hyperparams", {}) or {}
    cfg = PPOConfig.from_hyperparams(h, n_steps=args.n_steps, batch_size=args.batch_size, n_epochs=args.epochs)
    if "clip_range" not in h:
        cfg.clip_range = args.clip_range
    if "entrop_coef" not in h:
        cfg.ent_coef = args.ent_coef
    if "learning_rate" not in h:
        cfg.learning_rate = args.lr
    C = ns.num_communication
    if args.recurrent:
        def make_learner(seed):
            return RecurrentPPO(env.obs_width, 4, C, args.envs, env.device, cfg, seed=seed, lstm_hidden=args.lstm_hidden)
    else:
        def make_learner(seed):
            return PPO(env.obs_width, 4, C, args.envs, env.device, cfg, seed=seed)
    ego = make_learner(args.seed)
    partner = BatchedOnPolicyAgent(make_learner(args.seed + 1))
    penv = PantheonVecEnv(env, partner, reward_scale=args.reward_scale)

    if args.total_timesteps > 0:                            # whole rollouts, like OnPolicyAlgorithm.learn
        per_iter = args.n_steps * args.envs * world
        args.iters = max(1, -(-args.total_timesteps // per_iter))
    obs = penv.reset()
    starts = torch.ones(args.envs, device=env.device)
    # On a GPU with feed-forward learners the rollout is ONE CUDA graph (n_steps x [ego policy, partner policy, oc_step,
    # buffer writes]; the step kernel writes the observations straight into the rollout buffers' storage) and so is each
    # learner's minibatch update (ppo.PPO.train); `--no-graph` keeps the eager loop, which is also what the LSTM learner
    # and the CPU tests use.  The rollout has no collective in it, so it is graphed under torchrun too.
    cfg.cuda_graph = not args.no_graph
    graphed = (not args.no_graph) and (not args.recurrent) and env.device.type == "cuda"
    roll = GraphedRollout(penv, ego) if graphed else None
    t0 = time.time()
    steps = 0
    history = []
    for it in range(1, args.iters + 1):
        if roll is not None:
            obs, starts = roll.run()
            last_values = ego.value(obs, starts)
            ego.buffer.compute_returns_and_advantage(last_values, starts)
            stats = ego.train()
        else:
            obs, starts, stats = collect_and_train(penv, ego, obs, starts)
        steps += args.n_steps * args.envs * world
        if it % args.log_every == 0 or it == args.iters:
            ep = penv.pop_episode_stats()
            dt = time.time() - t0
            line = dict(iter=it, env_steps=steps, agent_steps_per_s=2 * steps / dt, wall_s=dt, world=world,
                        cuda_graphs=bool(graphed), **ep, ego_loss=stats, partner_updates=partner.iteration)
            history.append(line)
            if rank == 0:
                print(json.dumps(line), flush=True)
    if args.eval_steps > 0:
        # tester.py-style evaluation (tester.py:72-106): deterministic (argmax) actions, no learning
        obs = penv.reset()
        penv.pop_episode_stats()

        class _Frozen:                      # partner with the same policy, recording nothing
            starts = torch.ones(args.envs, device=env.device)

            def get_action(self, o, record=True):
                return partner.model.act(o, self.starts, deterministic=True)[0]

            def update(self, r, d):
                self.starts.copy_(d)
        penv.add_partner_agent(_Frozen())
        ego_starts = torch.ones(args.envs, device=env.device)
        for _ in range(args.eval_steps):
            obs, _, d = penv.step(ego.act(obs, ego_starts, deterministic=True)[0].to(torch.int32))
            ego_starts = d.to(torch.float32)
        ev = penv.pop_episode_stats()
        if rank == 0:
            print(json.dumps(dict(eval=True, steps=args.eval_steps, **ev)), flush=True)
        history.append(dict(eval=True, **ev))
    if args.save_dir and rank == 0:                        # the ranks' weights are identical
        os.makedirs(args.save_dir, exist_ok=True)
        save_learner(ego, os.path.join(args.save_dir, "ppo_ego.pt"))
        save_learner(partner.model, os.path.join(args.save_dir, "ppo_partner1.pt"))
    if learners_out is not None:
        learners_out += [ego, partner.model]
    env.close()
    return history


if __name__ == "__main__":
    sys.exit(0 if main() else 1)
