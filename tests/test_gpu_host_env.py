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


@pytest.mark.parametrize("level,A,T,C,E", [("open-divider_tomato", 2, 37, 10, 3001),
                                           ("random-salad-superwide", 2, 23, 100, 1111),
                                           ("partial-divider_salad", 3, 29, 6, 777)])
def test_host_env_compact_i8_format_vs_c_oracle(level, A, T, C, E):
    """oc_reset_host_i8 / oc_step_host_i8: int8 rows + per-env clock carry exactly the oracle's observation
    (integers compared as integers, the clock as float32), terminal rows and terminal clocks included."""
    from gym_comm_b200.host_env import OvercookedHostVecEnv
    cfg = dict(level=level, num_agents=A, max_num_timesteps=T, communication_on=True, num_communication=C,
               ego_led=False, fow_radius=2, ego_config=D, partner_config=D)
    text = levels_data.LEVELS[level]
    subtasks = levels_data.SUBTASKS[tuple(text.split("\n\n")[1].split("\n"))]
    env = OvercookedHostVecEnv(argparse.Namespace(**cfg), num_envs=E, seed=31, obs_format="i8")
    ora = COracle(text, subtasks, E, seed=31, **{k: v for k, v in cfg.items() if k != "level"})
    rng = np.random.default_rng(4)
    F = env.obs_width

    def same(i8, ts, ref):
        return (i8.dtype == np.int8 and i8.shape == (E, A, F - 1) and np.array_equal(i8, ref[..., :-1].astype(np.int64))
                and np.array_equal(ts, ref[:, 0, -1].astype(np.float32)))
    assert same(env.reset(), env.timestep, ora.reset())
    term_o = np.zeros((E, A, F))
    warm = (np.arange(E) % 7 == 0).astype(np.uint8)
    for t in range(2 * T + 9):
        a = np.stack([rng.integers(0, 4, (E, A)), rng.integers(0, C, (E, A))], -1).astype(np.int32)
        obs, rew, done, infos = env.step(a)
        oo, orr, od = ora.step(a, auto_reset=True, term_obs=term_o)
        assert np.array_equal(done, od.astype(bool)), t
        assert np.array_equal(rew[:, 0], orr.astype(np.float32)), t
        assert same(obs, env.timestep, oo), t
        for e in np.flatnonzero(done):
            assert np.array_equal(infos[e]["terminal_observation"], term_o[e][:, :-1].astype(np.int64)), (t, e)
            assert infos[e]["terminal_timestep"] == np.float32(term_o[e][0, -1]), (t, e)
        if t == 5:
            assert same(env.reset(mask=warm), env.timestep, ora.reset(mask=warm))
    assert np.array_equal(env.obs_float(), oo.astype(np.float32))
    d = env.obs_dict()
    assert d["timestep"].shape == (E, A, 1) and d["agent1_comm"].dtype == np.int8
    assert sum(v.shape[-1] for v in d.values()) == F
    env.close()
    ora.close()


def test_device_pack_obs_i8_matches_torch():
    """oc_pack_obs_i8 on device tensors (cfg5-wide rows, batch not a multiple of the CTA) against a torch slice."""
    import torch
    from gym_comm_b200.vec_env import OvercookedVecEnv
    E = 20011
    ns = argparse.Namespace(level="random-salad-superwide", num_agents=2, max_num_timesteps=17, communication_on=True,
                            num_communication=100, ego_led=False, fow_radius=2, ego_config=D, partner_config=D)
    env = OvercookedVecEnv(ns, num_envs=E, device="cuda:0", seed=5, auto_reset=True)
    n0 = env.launch_count()
    obs = torch.empty((6, E, 2, env.obs_width), device="cuda:0")
    env.rollout(6, obs_out=obs)
    for s in range(6):
        i8, ts = env.pack_obs_i8(obs[s])
        assert torch.equal(i8.to(torch.float32), obs[s][..., :-1]) and torch.equal(ts, obs[s][:, 0, -1])
    assert env.launch_count() - n0 == 7
    env.close()


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
