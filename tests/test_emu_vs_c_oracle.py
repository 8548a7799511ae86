"""The device code (tests/emu: oc_device.cuh compiled for the host, TEST ONLY) against the C oracle
on ragged multi-warp batches with auto-reset and terminal observations, under every shared-memory
row format the host can select (float rows of 32 / 16 / 8 envs per pass, padded or not, byte rows).
The GPU suite repeats this at BASELINE.json's sizes; this one runs without a GPU."""
import argparse

import numpy as np
import pytest
import torch

from gym_comm_b200 import levels_data
from gym_comm_b200.vec_env import OvercookedVecEnv
from oracle.c_oracle import COracle
from tests.parity_util import EmuVecEnv, emu_library

D = dict(CAN_MOVE=True, ALLERGIC=False, BLIND=False)
CONFIGS = {
    "tomato_c10": dict(level="open-divider_tomato", num_agents=2, max_num_timesteps=23, communication_on=True,
                       num_communication=10, ego_led=False, fow_radius=2, ego_config=D, partner_config=D),
    "salad3_c13": dict(level="partial-divider_salad", num_agents=3, max_num_timesteps=17, communication_on=True,
                       num_communication=13, ego_led=False, fow_radius=2, ego_config=D, partner_config=D),
    "random_wide_c100": dict(level="random-salad-superwide", num_agents=2, max_num_timesteps=19,
                             communication_on=True, num_communication=100, ego_led=False, fow_radius=2,
                             ego_config=D, partner_config=D),
    # rows of 96 floats (a multiple of 8 words): grouped rows -- four contiguous rows, then 16 bytes of padding
    "salad_c8_grouped": dict(level="random-open-divider_salad_small_cramped", num_agents=2, max_num_timesteps=19,
                             communication_on=True, num_communication=8, ego_led=False, fow_radius=10,
                             ego_config=dict(CAN_MOVE=False, ALLERGIC=True, BLIND=False),
                             partner_config=dict(CAN_MOVE=True, ALLERGIC=False, BLIND=True)),
    "wide3_c60": dict(level="partial-divider_salad", num_agents=3, max_num_timesteps=13, communication_on=True,
                      num_communication=60, ego_led=False, fow_radius=2, ego_config=D, partner_config=D),
    "wide4_c100": dict(level="open-divider_salad", num_agents=4, max_num_timesteps=11, communication_on=True,
                       num_communication=100, ego_led=True, fow_radius=1, ego_config=D, partner_config=D),
}
# a kitchen with one Tomato and one Plate: the (2 object slots, 1 food channel) kernel shape
TINY_TOMATO = "--*--\nt   p\n/   -\n-----\n\nSimpleTomato\n\n1 1\n3 1\n2 2\n1 2"
CONFIGS["tiny_tomato_shape21"] = dict(level="tiny", level_text=TINY_TOMATO, num_agents=2, max_num_timesteps=21,
                                      communication_on=True, num_communication=4, ego_led=False, fow_radius=1,
                                      ego_config=D, partner_config=D)
FORMATS = {
    "default": {},
    "f16": dict(OC_ROW_FORMAT="f", OC_ROW_ENVS="16"),
    "f8_nopad": dict(OC_ROW_FORMAT="f", OC_ROW_ENVS="8", OC_ROW_PAD="0"),
    "f4": dict(OC_ROW_FORMAT="f", OC_ROW_ENVS="4"),
    "bytes": dict(OC_ROW_FORMAT="b"),
    "f16x2": dict(OC_ROW_BUFS="2"),
    "f8x2": dict(OC_ROW_FORMAT="f", OC_ROW_ENVS="16", OC_ROW_PAD="0", OC_ROW_BUFS="2"),
    "group1": dict(OC_ROW_GROUP="1"),          # every row padded (one bulk copy per row on the device)
    "group8": dict(OC_ROW_GROUP="8"),
}


@pytest.mark.parametrize("fmt", sorted(FORMATS))
@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_step_autoreset_terminal_obs(name, fmt, monkeypatch):
    for k, v in FORMATS[fmt].items():
        monkeypatch.setenv(k, v)
    cfg = dict(CONFIGS[name])
    text = cfg.pop("level_text", None) or levels_data.LEVELS[cfg["level"]]
    subtasks = levels_data.SUBTASKS[tuple(text.split("\n\n")[1].split("\n"))]
    E, n, seed = 75, cfg["num_agents"], 99
    env = EmuVecEnv(argparse.Namespace(**cfg), num_envs=E, device="cpu", lib=emu_library(), seed=seed,
                           auto_reset=True, level_text=text)
    ora = COracle(text, subtasks, E, seed=seed, **{k: v for k, v in cfg.items() if k != "level"})
    rng = np.random.default_rng(5)
    term = torch.full((E, n, env.obs_width), -7.0)
    term_o = np.full((E, n, env.obs_width), -7.0)
    assert np.array_equal(env.reset().numpy(), ora.reset().astype(np.float32))
    for t in range(3 * cfg["max_num_timesteps"] + 2):
        a = np.stack([rng.integers(0, 4, (E, n)), rng.integers(0, cfg["num_communication"], (E, n))], -1).astype(np.int32)
        obs, rew, done = env.step(torch.from_numpy(a), term_obs_out=term, want_f64=True)
        oo, orr, od = ora.step(a, auto_reset=True, term_obs=term_o)
        assert np.array_equal(env.rewards64.numpy(), orr), (name, fmt, t)
        assert np.array_equal(done.numpy(), od), (name, fmt, t)
        assert np.array_equal(obs.numpy(), oo.astype(np.float32)), (name, fmt, t)
    assert np.array_equal(term.numpy(), term_o.astype(np.float32))
    assert env.decode_state()["episodes"].min() >= 3

    # the fused rollout continues from the same state with the shared Philox action spec, across an episode boundary
    # of every env (single-pass float rows are not cleared between the steps of a launch: what an env's row shows of
    # the finished episode -- message one-hots, completed subtasks -- has to be taken back)
    R = cfg["max_num_timesteps"] + 3
    obs = torch.zeros((R, E, n, env.obs_width))
    done = torch.zeros((R, E), dtype=torch.uint8)
    env.rollout(R, obs_out=obs, done_out=done)
    oo, orr, od, _ = ora.rollout(R, want_obs=True, want_actions=True)
    assert np.array_equal(done.numpy(), od) and np.array_equal(obs.numpy(), oo.astype(np.float32))
    assert done.numpy().any(axis=0).all()
    env.close()
    ora.close()
