"""Python side of the host-buffer env (`OvercookedHostVecEnv`: pinned buffers, the two observation formats,
SB3 infos / terminal observations, the word-wise scan for finished envs) on the CPU emulation of the device
functions, against the C oracle.  The CUDA entry points themselves are the GPU tests' job
(tests/test_gpu_host_env.py)."""
import argparse

import numpy as np
import pytest

from gym_comm_b200 import levels_data
from gym_comm_b200.host_env import OvercookedHostVecEnv, finished_indices
from oracle.c_oracle import COracle
from tests.parity_util import EmuHostLibrary

D = dict(CAN_MOVE=True, ALLERGIC=False, BLIND=False)


@pytest.mark.parametrize("E", [1, 7, 8, 64, 1000, 4099])
def test_finished_indices_equals_flatnonzero(E):
    rng = np.random.default_rng(E)
    buf = np.zeros((E + 7) // 8 * 8, np.uint8)
    for n in sorted({0, 1, min(E, 3), E // 50, E // 3, E}):
        buf[:] = 0
        buf[rng.choice(E, n, replace=False)] = 1
        got = finished_indices(buf, E)
        assert got == np.flatnonzero(buf[:E]).tolist() and all(type(e) is int for e in got)


@pytest.mark.parametrize("fmt", ["f32", "i8", "i8-separate-buffers"])
@pytest.mark.parametrize("level,A,T,C,E", [("open-divider_tomato", 2, 11, 5, 75), ("partial-divider_salad", 3, 9, 4, 41)])
def test_host_env_on_the_emulation_vs_c_oracle(fmt, level, A, T, C, E):
    """"i8" runs the one-block path (`oc_step_host_block`: u8 actions, per-env reward, compact rows from the step
    kernel); "i8-separate-buffers" the older `oc_step_host_i8` entry points a handle falls back to when its rows are
    too wide for the compact kernels (float rows + repack + gather)."""
    compact = fmt != "i8-separate-buffers"
    fmt = "i8" if fmt.startswith("i8") else fmt
    cfg = dict(level=level, num_agents=A, max_num_timesteps=T, communication_on=True, num_communication=C,
               ego_led=False, fow_radius=2, ego_config=D, partner_config=D)
    text = levels_data.LEVELS[level]
    subtasks = levels_data.SUBTASKS[tuple(text.split("\n\n")[1].split("\n"))]
    lib = EmuHostLibrary(compact=compact)
    env = OvercookedHostVecEnv(argparse.Namespace(**cfg), num_envs=E, seed=31, obs_format=fmt, lib=lib)
    assert env._block_mode == (fmt == "i8" and compact)
    assert env.action_dtype == (np.uint8 if env._block_mode else np.int32)
    lib.bind(env)
    ora = COracle(text, subtasks, E, seed=31, **{k: v for k, v in cfg.items() if k != "level"})
    F = env.obs_width
    rng = np.random.default_rng(4)
    env.reset()
    assert np.array_equal(env.obs_float(), ora.reset().astype(np.float32))
    term_o = np.zeros((E, A, F))
    warm = (np.arange(E) % 3 == 0).astype(np.uint8)
    seen = 0
    for t in range(3 * T + 2):
        a = np.stack([rng.integers(0, 4, (E, A)), rng.integers(0, C, (E, A))], -1)          # int64: converted by step()
        obs, rew, done, infos = env.step(a)
        oo, orr, od = ora.step(a.astype(np.int32), auto_reset=True, term_obs=term_o)
        assert np.array_equal(done, od.astype(bool)) and np.array_equal(rew[:, 0], orr.astype(np.float32)), t
        assert np.array_equal(env.obs_float(), oo.astype(np.float32)), t
        assert obs.dtype == (np.int8 if fmt == "i8" else np.float32) and obs.shape == (E, A, F - (fmt == "i8"))
        for e in range(E):                                  # infos: exactly the finished envs carry terminal rows
            if not done[e]:
                assert infos[e] == {}, (t, e)
                continue
            seen += 1
            ref = term_o[e].astype(np.float32)
            if fmt == "i8":
                assert np.array_equal(infos[e]["terminal_observation"], ref[:, :-1]) and infos[e]["terminal_timestep"] == ref[0, -1]
            else:
                assert np.array_equal(infos[e]["terminal_observation"], ref)
        if t == 4:                                          # masked reset staggers the clocks: few envs finish per step later
            env.reset(mask=warm)
            assert np.array_equal(env.obs_float(), ora.reset(mask=warm).astype(np.float32))
    assert seen >= 2 * E
    d = env.obs_dict()
    assert d["timestep"].shape == (E, A, 1) and sum(v.shape[-1] for v in d.values()) == F
    assert np.array_equal(np.concatenate([d[k] for k in sorted(d)], -1).astype(np.float32), env.obs_float())
    with pytest.raises(ValueError):
        env.step(np.zeros((E, A), np.int32))
    env.close()
    env.close()
    ora.close()
