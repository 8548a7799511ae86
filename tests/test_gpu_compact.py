"""GPU: the compact-row kernels (`oc_step_i8` / `oc_reset_i8`, kernel MODE 3) against the float rows of a twin env,
ragged batch sizes, u8 actions, per-env reward, terminal rows; `oc_set_state` sanitising; the older host entry
point `oc_step_host_i8` with PAGEABLE caller buffers (staged copies + device-side gather of the terminal rows)."""
import argparse
import ctypes as C

import numpy as np
import pytest

from tests import compact_cases as cases

pytestmark = pytest.mark.gpu


def _make(ns, **kw):
    from gym_comm_b200.vec_env import OvercookedVecEnv
    return OvercookedVecEnv(ns, device="cuda:0", **kw)


@pytest.mark.parametrize("level,A,T,C,fow", cases.CASES)
@pytest.mark.parametrize("E,u8,per_env", [(4099, False, False), (33, True, True), (1, True, False), (20000, True, True)])
def test_step_i8_equals_float_rows(level, A, T, C, fow, E, u8, per_env):
    cases.run_step_i8_equals_float_rows(_make, "cuda:0", level, A, T, C, fow, E, u8, per_env)


def test_step_i8_full_size_crosses_episode_boundaries():
    """BASELINE configs[1] batch (65,536 envs), staggered clocks: every env finishes within the run."""
    cases.run_step_i8_equals_float_rows(_make, "cuda:0", "open-divider_tomato", 2, 40, 10, 2, 65536, True, True, steps=90)


def test_set_state_sanitises():
    cases.run_set_state_sanitises(_make, "cuda:0")


def test_step_host_i8_with_pageable_buffers_vs_c_oracle():
    """`oc_step_host_i8` straight through ctypes with ordinary numpy arrays (not page-locked): the staged path with
    its device-side gather of the compact terminal rows."""
    from gym_comm_b200 import _cabi, levels_data
    from gym_comm_b200.host_env import OvercookedHostVecEnv
    from oracle.c_oracle import COracle
    level, A, T, Cc, E = "open-divider_salad", 2, 19, 9, 1500
    cfg = dict(level=level, num_agents=A, max_num_timesteps=T, communication_on=True, num_communication=Cc,
               ego_led=False, fow_radius=2, ego_config=cases.D, partner_config=cases.D)
    text = levels_data.LEVELS[level]
    subtasks = levels_data.SUBTASKS[tuple(text.split("\n\n")[1].split("\n"))]
    env = OvercookedHostVecEnv(argparse.Namespace(**cfg), num_envs=E, seed=31, obs_format="i8")
    ora = COracle(text, subtasks, E, seed=31, **{k: v for k, v in cfg.items() if k != "level"})
    lib, h, F = env.lib, env._handle, env.obs_width
    p = lambda a: C.c_void_p(a.ctypes.data)
    obs8, ts = np.zeros((E, A, F - 1), np.int8), np.zeros(E, np.float32)
    rew, done = np.zeros((E, A), np.float32), np.zeros(E, np.uint8)
    term8, term_ts = np.zeros((E, A, F - 1), np.int8), np.zeros(E, np.float32)
    lib.check(lib.reset_host_i8(h, None, None, p(obs8), p(ts), None), "oc_reset_host_i8")
    ref = ora.reset()
    assert np.array_equal(obs8, ref[..., :-1].astype(np.int64)) and np.array_equal(ts, ref[:, 0, -1].astype(np.float32))
    env.stagger_clocks(7)                       # through oc_get_state_host / oc_set_state_host
    st = env.get_state()
    ora_t = (np.arange(E) % 7).astype(np.uint32)
    assert np.array_equal(st[:, 0] & 0xFFFF, ora_t)
    ora.set_clocks(ora_t)
    term_o = np.zeros((E, A, F))
    rng = np.random.default_rng(8)
    for t in range(2 * T + 3):
        a = np.stack([rng.integers(0, 4, (E, A)), rng.integers(0, Cc, (E, A))], -1).astype(np.int32)
        lib.check(lib.step_host_i8(h, p(a), p(obs8), p(ts), p(rew), None, p(done), p(term8), p(term_ts),
                                   _cabi.OC_FLAG_AUTO_RESET, None), "oc_step_host_i8")
        oo, orr, od = ora.step(a, auto_reset=True, term_obs=term_o)
        assert np.array_equal(done, od), t
        assert np.array_equal(rew[:, 0], orr.astype(np.float32)), t
        assert np.array_equal(obs8, oo[..., :-1].astype(np.int64)) and np.array_equal(ts, oo[:, 0, -1].astype(np.float32)), t
        d = done.astype(bool)
        assert np.array_equal(term8[d], term_o[d][..., :-1].astype(np.int64)), t
        assert np.array_equal(term_ts[d], term_o[d][:, 0, -1].astype(np.float32)), t
    env.close()
    ora.close()
