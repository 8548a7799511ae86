"""Small end-to-end exercise of every kernel for `compute-sanitizer --tool memcheck|racecheck`."""
import argparse
import sys

import torch

sys.path.insert(0, ".")
from gym_comm_b200.vec_env import OvercookedVecEnv  # noqa: E402

d = dict(CAN_MOVE=True, ALLERGIC=False, BLIND=False)
CASES = [
    dict(level="open-divider_tomato", num_agents=2, max_num_timesteps=20, num_communication=10, fow_radius=2),
    dict(level="partial-divider_salad", num_agents=3, max_num_timesteps=15, num_communication=10, fow_radius=2),
    dict(level="random-open-divider_salad_small_cramped", num_agents=2, max_num_timesteps=12, num_communication=8, fow_radius=10),
    dict(level="random-salad-superwide", num_agents=2, max_num_timesteps=10, num_communication=100, fow_radius=2),
    dict(level="open-divider_tl", num_agents=3, max_num_timesteps=10, num_communication=5, fow_radius=3),
]
for c in CASES:
    ns = argparse.Namespace(communication_on=True, ego_led=False, ego_config=d, partner_config=d, **c)
    for E in (70, 300):
        env = OvercookedVecEnv(ns, num_envs=E, device="cuda:0", seed=3)
        A, F = env.num_agents, env.obs_width
        g = torch.Generator(device="cuda:0").manual_seed(0)
        term = torch.zeros((E, A, F), device="cuda:0")
        for t in range(30):
            a = torch.stack([torch.randint(0, 4, (E, A), generator=g, device="cuda:0", dtype=torch.int32),
                             torch.randint(0, c["num_communication"], (E, A), generator=g, device="cuda:0", dtype=torch.int32)], -1).contiguous()
            env.step(a, term_obs_out=term, want_f64=True)
        obs = torch.zeros((8, E, A, F), device="cuda:0")
        rew = torch.zeros((8, E, A), device="cuda:0")
        done = torch.zeros((8, E), dtype=torch.uint8, device="cuda:0")
        acts = torch.zeros((8, E, A, 2), dtype=torch.int32, device="cuda:0")
        for _ in range(3):
            env.rollout(8, obs_out=obs, rew_out=rew, done_out=done, actions_out=acts)
        env.replay(acts, obs_out=obs, rew_out=rew, done_out=done)
        env.reset(mask=done[-1].contiguous())
        env.set_state(env.get_state())
        env.stats()
        torch.cuda.synchronize()
        env.close()
print("sanitize_case ok")
