"""Where the fused rollout's time goes: the same launch with all outputs, without observations
(dynamics + reward/done only) and with observations only.  usage (GPU box):
python tools/rollout_breakdown.py [workload]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, workload_namespace  # noqa: E402
from gym_comm_b200.vec_env import OvercookedVecEnv  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    w = WORKLOADS[name]
    E, R, N = w["envs"], 64, 12
    dev = torch.device("cuda", 0)
    env = OvercookedVecEnv(workload_namespace(w), num_envs=E, device=dev, seed=1)
    A, F = env.num_agents, env.obs_width
    obs = torch.empty((R, E, A, F), device=dev)
    rew = torch.empty((R, E, A), device=dev)
    done = torch.empty((R, E), dtype=torch.uint8, device=dev)
    variants = {"all outputs": dict(obs_out=obs, rew_out=rew, done_out=done),
                "no observations": dict(rew_out=rew, done_out=done),
                "observations only": dict(obs_out=obs),
                "no outputs": dict()}
    for label, kw in variants.items():
        for _ in range(3):
            env.rollout(R, **kw)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(N):
            env.rollout(R, **kw)
        b.record()
        torch.cuda.synchronize()
        print("%s %-18s %.2f us/step" % (name, label, a.elapsed_time(b) * 1e3 / (N * R)))


if __name__ == "__main__":
    main()
