#!/usr/bin/env python
"""Evaluate saved ego / partner policies on the GPU env -- the counterpart of the reference's tester.py
(`tester.py:72-128`: load both agents as fixed policies, play `--total-episodes` games, print every
episode's total reward, then the average and the standard deviation).

    python train_ppo.py --save-dir model/tomato
    python evaluate_policy.py -j env_args.json --ego-load model/tomato/ppo_ego.pt --alt-load model/tomato/ppo_partner1.pt -t 10000

The games run `--envs` at a time on the device (`OvercookedVecEnv` + `PantheonVecEnv` with a
`BatchedStaticPolicyAgent` partner).  Every env plays a fixed QUOTA of games (N // E, the first N % E envs one
more) and only those count: auto-resetting envs that stopped at "the first N finished games" would over-count
short (successful) episodes, whose envs replay while the long ones are still running -- the reference plays its N
games one after the other and has no such bias.  Like the reference's `StaticPolicyAgent` both agents sample from their
policies; `--deterministic` takes the argmax instead.  `--render` prints env 0's ASCII display every step
(`env.render()`, tester.py:79-91).  The last line is one JSON object with the statistics.
"""
import argparse
import json
import sys

import torch

from gym_comm_b200 import OvercookedVecEnv, create_arglist, namespace_from_dict
from gym_comm_b200.pantheon import BatchedStaticPolicyAgent, PantheonVecEnv
from gym_comm_b200.ppo import load_learner


def main(argv=None, env_factory=None):
    """`env_factory(ns, args)` lets the CPU tests put the emulated env under the same loop."""
    ap = argparse.ArgumentParser(description="play saved ego / partner policies on the batched GPU env")
    ap.add_argument("--json-path", "-j", default=None, help="env config JSON (tester.py --json-path)")
    ap.add_argument("--level", default="open-divider_tomato", help="used when no --json-path is given")
    ap.add_argument("--max-num-timesteps", type=int, default=200)
    ap.add_argument("--num-communication", type=int, default=10)
    ap.add_argument("--ego-load", required=True, help="ego policy file written by train_ppo.py --save-dir (ppo_ego.pt)")
    ap.add_argument("--alt-load", required=True, help="partner policy file (ppo_partner1.pt)")
    ap.add_argument("--total-episodes", "-t", type=int, default=100, help="stop after this many finished games")
    ap.add_argument("--envs", type=int, default=0, help="games played at once (default: min(total episodes, 65536))")
    ap.add_argument("--device", "-d", default="cuda:0")
    ap.add_argument("--deterministic", action="store_true")
    ap.add_argument("--render", action="store_true", help="print env 0 as it is being run")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--max-steps", type=int, default=0, help="safety stop after this many env steps (0 = none)")
    args = ap.parse_args(argv)

    if args.json_path:
        ns = create_arglist(args.json_path)
    else:
        ns = namespace_from_dict(dict(level=args.level, num_agents=2, max_num_timesteps=args.max_num_timesteps,
                                      communication_on=True, num_communication=args.num_communication))
    args.envs = args.envs or max(1, min(args.total_episodes, 65536))
    torch.manual_seed(args.seed)
    env = env_factory(ns, args) if env_factory is not None else \
        OvercookedVecEnv(ns, num_envs=args.envs, device=args.device, seed=args.seed, auto_reset=True)
    ego = load_learner(args.ego_load, args.envs, env.device)
    alt = load_learner(args.alt_load, args.envs, env.device)
    for name, m in (("ego", ego), ("alt", alt)):
        if m.buffer.obs.shape[-1] != env.obs_width or m.policy.num_comm != ns.num_communication:
            raise ValueError("%s policy was trained for another observation / message width" % name)
    penv = PantheonVecEnv(env, BatchedStaticPolicyAgent(alt, args.deterministic))

    obs = penv.reset()
    E, dev = args.envs, env.device
    starts = torch.ones(E, device=dev)
    # per-env quota of games: exactly --total-episodes in all
    quota = torch.full((E,), args.total_episodes // E, device=dev, dtype=torch.int64)
    quota[:args.total_episodes % E] += 1
    played = torch.zeros(E, device=dev, dtype=torch.int64)
    ret = torch.zeros(E, device=dev, dtype=torch.float64)
    length = torch.zeros(E, device=dev, dtype=torch.float64)
    acc = torch.zeros(5, device=dev, dtype=torch.float64)          # games, sum return, sum return^2, sum length, delivered
    T = float(ns.max_num_timesteps)
    steps = 0
    if args.render:
        print(env.render(0))
    while not (args.max_steps and steps >= args.max_steps):
        act = ego.act(obs, starts, deterministic=args.deterministic)[0]
        obs, rew, done = penv.step(act.to(torch.int32))
        starts = done.to(torch.float32)
        steps += 1
        ret += rew
        length += 1
        d = done.bool()
        m = d & (played < quota)                                     # this game is one of the env's quota
        acc += torch.stack([m.sum(), (ret * m).sum(), (ret * ret * m).sum(), (length * m).sum(),
                            (m & (length < T)).sum()]).to(torch.float64)
        played += m
        ret.masked_fill_(d, 0.0)
        length.masked_fill_(d, 0.0)
        if args.render:
            print(env.render(0))
        if (steps % 16 == 0 or args.render) and bool((played >= quota).all()):   # one host sync every 16 steps
            break
    games, rsum, rsq, lsum, deliv = acc.tolist()
    n = max(games, 1.0)
    mean = rsum / n
    out = dict(episodes=int(games), env_steps=steps * args.envs, average_reward=mean,
               standard_deviation=max(rsq / n - mean * mean, 0.0) ** 0.5,
               ep_len_mean=lsum / n, delivered_frac=deliv / n)
    print("Average Reward: ", out["average_reward"])
    print("Standard Deviation: ", out["standard_deviation"])
    print(json.dumps(out), flush=True)
    env.close()
    return out


if __name__ == "__main__":
    sys.exit(0 if main() else 1)
