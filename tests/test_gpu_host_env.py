"""GPU: the host-buffer path (oc_reset_host / oc_step_host behind `OvercookedHostVecEnv`, numpy only)
against the C oracle: observations, rewards, dones and SB3-style terminal observations, on the
branch where a few envs finish in a step (row-wise copies) and on the one where all of them do."""
import argparse
import subprocess
import sys
import os

import numpy as np
import pytest

from gym_comm_b200 import levels_data
from oracle.c_oracle import COracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = dict(CAN_MOVE=True, ALLERGIC=False, BLIND=False)


@pytest.mark.parametrize("level,A,T,C,E", [("open-divider_tomato", 2, 37, 10, 3000),
                                           ("random-salad-superwide", 2, 23, 100, 1111),
                                           ("partial-divider_salad", 3, 29, 6, 777)])
def test_host_env_vs_c_oracle(level, A, T, C, E):
    from gym_comm_b200.host_env import OvercookedHostVecEnv
    cfg = dict(level=level, num_agents=A, max_num_timesteps=T, communication_on=True, num_communication=C,
               ego_led=False, fow_radius=2, ego_config=D, partner_config=D)
    text = levels_data.LEVELS[level]
    subtasks = levels_data.SUBTASKS[tuple(text.split("\n\n")[1].split("\n"))]
    env = OvercookedHostVecEnv(argparse.Namespace(**cfg), num_envs=E, seed=31)
    ora = COracle(text, subtasks, E, seed=31, **{k: v for k, v in cfg.items() if k != "level"})
    rng = np.random.default_rng(3)
    assert np.array_equal(env.reset(), ora.reset().astype(np.float32))
    term_o = np.zeros((E, A, env.obs_width))
    # stagger the episode clocks so that later steps finish only a few envs at a time
    warm = (np.arange(E) % 7 == 0).astype(np.uint8)
    for t in range(2 * T + 9):
        a = np.stack([rng.integers(0, 4, (E, A)), rng.integers(0, C, (E, A))], -1).astype(np.int32)
        obs, rew, done, infos = env.step(a)
        oo, orr, od = ora.step(a, auto_reset=True, term_obs=term_o)
        assert np.array_equal(done, od.astype(bool)), t
        assert np.array_equal(rew[:, 0], orr.astype(np.float32)), t
        assert np.array_equal(obs, oo.astype(np.float32)), t
        for e in np.flatnonzero(done):
            assert np.array_equal(infos[e]["terminal_observation"], term_o[e].astype(np.float32)), (t, e)
        if t == 5:                                   # masked reset through the host path
            assert np.array_equal(env.reset(mask=warm), ora.reset(mask=warm).astype(np.float32))
    assert np.array_equal(env.terminal_obs, term_o.astype(np.float32))
    views = env.obs_dict()
    assert views["timestep"].shape == (E, A, 1)
    env.close()
    ora.close()


def test_host_env_does_not_import_torch():
    code = ("import sys, argparse, numpy as np\n"
            "from gym_comm_b200 import OvercookedHostVecEnv\n"
            "ns = argparse.Namespace(level='open-divider_tomato', num_agents=2, max_num_timesteps=20)\n"
            "env = OvercookedHostVecEnv(ns, num_envs=100)\n"
            "env.reset()\n"
            "n = 0\n"
            "for t in range(45):\n"
            "    a = np.zeros((100, 2, 2), dtype=np.int64); a[..., 0] = t % 4\n"
            "    obs, rew, done, infos = env.step(a)\n"
            "    n += int(done.sum())\n"
            "assert n == 200, n\n"
            "assert 'torch' not in sys.modules\n"
            "env.close(); print('ok')\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr
