"""Every env config the reference ships (`env_args.json`, `spread/*.json`: 17 files) goes through the drop-in's
config front end, compiles, and steps on the CPU emulation of the device code in lock-step with the C oracle --
including the ten files the reference itself cannot start from (missing `ego_config` / `partner_config` /
`CAN_MOVE`), which get the defaults.  Needs /root/reference for the files (they are not copied into
the repo); the one config BASELINE.json names is also committed under configs/."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from gym_comm_b200 import create_arglist, levels_data
from gym_comm_b200.vec_env import OvercookedVecEnv
from oracle.c_oracle import COracle
from tests.parity_util import EmuVecEnv, emu_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
FILES = sorted(glob.glob(os.path.join(REF, "spread", "*.json")) + glob.glob(os.path.join(REF, "env_args.json")) +
               glob.glob(os.path.join(ROOT, "configs", "*.json")))


@pytest.mark.skipif(not FILES, reason="no config files found")
@pytest.mark.parametrize("path", FILES, ids=[os.path.relpath(p, REF if p.startswith(REF) else ROOT) for p in FILES])
def test_config_compiles_and_steps_like_the_oracle(path):
    raw = json.load(open(path))
    ns = create_arglist(path)
    assert ns.level == raw["level"] and ns.num_agents == raw["num_agents"]
    assert ns.num_communication == raw.get("num_communication", 10) and ns.fow_radius == raw.get("fow_radius", 2)
    for side in ("ego_config", "partner_config"):                  # forgotten keys -> defaults, given keys kept
        want = dict(CAN_MOVE=True, ALLERGIC=False, BLIND=False)
        want.update(raw.get(side, {}))
        assert getattr(ns, side) == want
    E = 37
    env = EmuVecEnv(ns, num_envs=E, device="cpu", seed=11, auto_reset=True, lib=emu_library())
    assert env.obs_width == 23 + len(env.level.subtasks) + 2 * ns.num_communication
    text = levels_data.LEVELS[ns.level]
    subtasks = levels_data.SUBTASKS[tuple(text.split("\n\n")[1].split("\n"))]
    ora = COracle(text, subtasks, E, seed=11, num_agents=ns.num_agents, max_num_timesteps=ns.max_num_timesteps,
                  communication_on=ns.communication_on, num_communication=ns.num_communication, ego_led=ns.ego_led,
                  fow_radius=ns.fow_radius, ego_config=ns.ego_config, partner_config=ns.partner_config)
    assert np.array_equal(env.reset().numpy(), ora.reset().astype(np.float32))
    rng = np.random.default_rng(5)
    A, C = ns.num_agents, ns.num_communication
    for t in range(40):
        a = np.stack([rng.integers(0, 4, (E, A)), rng.integers(0, C, (E, A))], -1).astype(np.int32)
        obs, rew, done = env.step(torch.from_numpy(a), want_f64=True)
        oo, orr, od = ora.step(a, auto_reset=True)
        assert np.array_equal(obs.numpy(), oo.astype(np.float32)), t
        assert np.array_equal(env.rewards64.numpy(), orr) and np.array_equal(done.numpy(), od), t
    env.close()
    ora.close()


REF_JSON = [p for p in FILES if p.startswith(REF)]


@pytest.mark.skipif(not REF_JSON, reason="needs /root/reference")
def test_namespace_equals_the_reference_parser_where_it_can_parse():
    """For the config files the reference's own `arglist.create_arglist` (arglist.py:96-121) accepts, our parser
    yields the same Namespace values for every key the env reads; the files it rejects fail there with a KeyError on
    `ego_config` / `partner_config` -- those are the ones that get our defaults."""
    import contextlib
    import importlib.util
    import io
    spec = importlib.util.spec_from_file_location("_ref_arglist", os.path.join(REF, "arglist.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    parsed, rejected = 0, 0
    for path in REF_JSON:
        ours = create_arglist(path)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                theirs = ref.create_arglist(path)
        except KeyError as ex:
            assert ex.args[0] in ("ego_config", "partner_config"), (path, ex)
            rejected += 1
            continue
        parsed += 1
        for key in ("level", "num_agents", "max_num_timesteps", "hyperparams", "communication_on", "num_communication",
                    "ego_led", "fow_radius", "total_timesteps", "record_interval", "wandb", "max_num_subtasks", "seed"):
            assert getattr(ours, key) == getattr(theirs, key), (path, key)
        for side in ("ego_config", "partner_config"):              # the reference keeps exactly the file's keys
            given = getattr(theirs, side)
            assert {k: getattr(ours, side)[k] for k in given} == given, (path, side)
    assert parsed >= 7 and parsed + rejected == len(REF_JSON)
