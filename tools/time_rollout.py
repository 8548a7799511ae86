"""Device time per fused step of an arbitrary config (GPU box): python tools/time_rollout.py level A C T fow E [blind1]"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_comm_b200 import namespace_from_dict
from gym_comm_b200.vec_env import OvercookedVecEnv

level, A, C, T, fow, E = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
blind1 = len(sys.argv) > 7 and sys.argv[7] == "blind1"
D = dict(CAN_MOVE=True, ALLERGIC=False, BLIND=False)
ns = namespace_from_dict(dict(level=level, num_agents=A, max_num_timesteps=T, communication_on=True, num_communication=C,
                              fow_radius=fow, ego_config=D, partner_config=dict(D, BLIND=blind1)))
dev = torch.device("cuda", 0)
env = OvercookedVecEnv(ns, num_envs=E, device=dev, seed=3, auto_reset=True)
n = 20
obs = torch.empty((64, E, A, env.obs_width), device=dev)
rew = torch.empty((64, E, A), device=dev)
done = torch.empty((64, E), dtype=torch.uint8, device=dev)
env.reset()
for _ in range(10):
    env.rollout(n, obs_out=obs[:n], rew_out=rew[:n], done_out=done[:n])
torch.cuda.synchronize()
ts = []
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(60):
    torch.cuda._sleep(200000)
    a.record()
    env.rollout(n, obs_out=obs[(rep % 3) * n:(rep % 3) * n + n], rew_out=rew[:n], done_out=done[:n])
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
print("%s A=%d C=%d F=%d blind1=%s: %.3f us/step (median of 60 x %d steps)" % (level, A, C, env.obs_width, blind1, ts[len(ts) // 2] * 1e3 / n, n))
