"""Where a training iteration of train_ppo.py spends its time (GPU box): rollout (CUDA graph), GAE, ego update, partner
update -- each bracketed by a device synchronisation.   python tools/train_profile.py [--batch-size N] [--epochs K]"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_comm_b200 import OvercookedVecEnv, namespace_from_dict  # noqa: E402
from gym_comm_b200.pantheon import BatchedOnPolicyAgent, GraphedRollout, PantheonVecEnv  # noqa: E402
from gym_comm_b200.ppo import PPO, PPOConfig  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--n-steps", type=int, default=32)
    ap.add_argument("--batch-size", type=int, default=65536)
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--iters", type=int, default=8)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    ns = namespace_from_dict(dict(level="open-divider_tomato", num_agents=2, max_num_timesteps=200, communication_on=True,
                                  num_communication=10))
    env = OvercookedVecEnv(ns, num_envs=a.envs, device=dev, seed=0, auto_reset=True)
    cfg = PPOConfig(n_steps=a.n_steps, batch_size=a.batch_size, n_epochs=a.epochs, learning_rate=1e-3, clip_range=0.2)
    ego = PPO(env.obs_width, 4, 10, a.envs, dev, cfg, seed=0)
    partner = BatchedOnPolicyAgent(PPO(env.obs_width, 4, 10, a.envs, dev, cfg, seed=1))
    penv = PantheonVecEnv(env, partner, reward_scale=0.1)
    penv.reset()
    roll = GraphedRollout(penv, ego)
    T = dict(partner_update=0.0, rollout=0.0, gae=0.0, ego_update=0.0)

    def tick():
        torch.cuda.synchronize()
        return time.perf_counter()
    for it in range(a.iters):
        timed = it >= 3                                   # eager warm-up, capture, first replays
        t0 = tick()
        partner.maybe_train()
        t1 = tick()
        obs, starts = roll.run()
        t2 = tick()
        ego.buffer.compute_returns_and_advantage(ego.value(obs, starts), starts)
        t3 = tick()
        ego.train()
        t4 = tick()
        if timed:
            T["partner_update"] += t1 - t0; T["rollout"] += t2 - t1; T["gae"] += t3 - t2; T["ego_update"] += t4 - t3
    n = a.iters - 3
    tot = sum(T.values()) / n
    print("envs %d n_steps %d batch %d epochs %d: %.1f ms / iteration = %.1f M agent-steps/s" %
          (a.envs, a.n_steps, a.batch_size, a.epochs, tot * 1e3, 2 * a.envs * a.n_steps / tot / 1e6))
    for k, v in T.items():
        print("   %-16s %7.1f ms  (%.0f %%)" % (k, v / n * 1e3, 100 * v / n / tot))
    nmb = a.epochs * (a.envs * a.n_steps // min(a.batch_size, a.envs * a.n_steps))
    print("   one graphed minibatch update: %.3f ms;  one rollout step (2 policies + oc_step + bookkeeping): %.3f ms" %
          (T["ego_update"] / n / nmb * 1e3, T["rollout"] / n / a.n_steps * 1e3))


if __name__ == "__main__":
    main()
