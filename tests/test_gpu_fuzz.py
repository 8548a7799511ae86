"""GPU: random kitchens (oracle/level_fuzz.py) -- the real kernels against the C oracle on geometry
the shipped levels never exercise, 512 envs each, uniform random actions, device-side auto-reset."""
import argparse

import numpy as np
import pytest
import torch

from gym_comm_b200 import levels_data
from gym_comm_b200.vec_env import OvercookedVecEnv
from oracle.c_oracle import COracle
from oracle.level_fuzz import random_level

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("seed", range(16))
def test_random_kitchen_on_gpu(seed):
    n = min(4, 2 + (seed % 3 == 1) + 2 * (seed % 8 == 7))
    text = random_level(5000 + seed, n)
    recipes = tuple(text.split("\n\n")[1].split("\n"))
    subtasks = levels_data.SUBTASKS[recipes]
    d = dict(CAN_MOVE=True, ALLERGIC=False, BLIND=False)
    cfg = dict(num_agents=n, max_num_timesteps=40 + 7 * (seed % 5), communication_on=(seed % 4 != 3),
               num_communication=2 + seed % 9, ego_led=(seed % 5 == 2), fow_radius=seed % 5,
               ego_config=dict(d, BLIND=(seed % 6 == 4)), partner_config=dict(d, ALLERGIC=(seed % 7 == 3)))
    E, T = 512 + seed, 260
    env = OvercookedVecEnv(argparse.Namespace(level="fuzz", **cfg), num_envs=E, device=DEV, seed=seed,
                           auto_reset=True, level_text=text, subtasks=subtasks)
    ora = COracle(text, subtasks, E, seed=seed, **cfg)
    gen = torch.Generator(device=DEV).manual_seed(seed)
    C = cfg["num_communication"]
    # bias the walk so objects get picked up, carried and put down: repeat the previous nav half the time
    prev = torch.randint(0, 4, (E, n), generator=gen, device=DEV, dtype=torch.int32)
    for t in range(T):
        fresh = torch.randint(0, 4, (E, n), generator=gen, device=DEV, dtype=torch.int32)
        keep = torch.rand((E, n), generator=gen, device=DEV) < 0.5
        prev = torch.where(keep, prev, fresh)
        a = torch.stack([prev, torch.randint(0, C, (E, n), generator=gen, device=DEV, dtype=torch.int32)], -1).contiguous()
        obs, rew, done = env.step(a, want_f64=True)
        oo, orr, od = ora.step(a.cpu().numpy(), auto_reset=True)
        assert torch.equal(env.rewards64.cpu(), torch.from_numpy(orr)), (seed, t)
        assert torch.equal(done.cpu(), torch.from_numpy(od)), (seed, t)
        assert torch.equal(obs.cpu(), torch.from_numpy(oo.astype(np.float32))), (seed, t)
    assert np.array_equal(env.decode_state()["episodes"], ora.state()["episodes"])
    env.close()
    ora.close()
