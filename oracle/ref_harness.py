"""TEST INFRASTRUCTURE ONLY -- loader for the *live* reference (kyle-he/gym-comm).

This module makes the unmodified reference env importable in the build
container, where ``gym``, ``termcolor``, ``matplotlib`` and
``stable_baselines3`` are not installed, by registering attribute-holder stub
modules before the import (recipe: SURVEY.md section 8c).  It is used by

* ``oracle/record_golden.py``  (generates ``tests/golden/*.npz``), and
* ``tests/test_oracle_vs_reference.py`` (skipped when no reference tree is
  reachable), and
* ``oracle/time_reference.py`` (the CPU arm of ``bench.py``).

Nothing here is imported by the product package.  The reference is imported
from where it lies (``/root/reference``); on the GPU box, where that path does
not exist, from the unmodified copy ``oracle/stage_ref.py`` staged into the
git-ignored ``oracle/_ref/``.

Facts the harness depends on (reference file:line):
* level path is cwd-relative      gym_cooking/envs/overcooked_environment.py:103
* ``world.py`` imports ``navigation_planner`` top-level  gym_cooking/utils/world.py:10
* subtask order depends on PYTHONHASHSEED               recipe_planner/stripsworld.py:72-77
* the env prints on reward events                       overcooked_environment.py:405-428
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

def _find_reference_root() -> str:
    """`/root/reference` in the build container; on the GPU box the byte-identical copy that
    `oracle/stage_ref.py` left in the git-ignored `oracle/_ref/` (it travels with gpurun)."""
    env = os.environ.get("OC_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/gym_cooking"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gym_cooking"))


def hashseed_is_canonical() -> bool:
    """Parity runs need PYTHONHASHSEED=0."""
    return os.environ.get("PYTHONHASHSEED") == "0"


class _Space:
    def __init__(self, *a, **k):
        self.args = a
        self.kwargs = k
        self.shape = k.get("shape")


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


def _install_stubs():
    if "gym" in sys.modules and getattr(sys.modules["gym"], "_oc_stub", False):
        return

    class Env:
        pass

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env

    gym = _mod("gym", Env=Env, Wrapper=Wrapper, _oc_stub=True)
    _mod("gym.error")
    _mod("gym.utils")
    _mod("gym.utils.seeding")
    spaces = _mod("gym.spaces")
    for n in ["Box", "Discrete", "MultiBinary", "MultiDiscrete", "Dict", "Tuple", "Space"]:
        setattr(spaces, n, type(n, (_Space,), {}))
    _mod("gym.envs")
    _mod("gym.envs.registration", register=lambda **k: None)
    gym.make = None

    _mod("termcolor", colored=lambda s, *a, **k: s)
    _mod("matplotlib")
    _mod("matplotlib.pyplot")

    class _Dummy:
        def __init__(self, *a, **k):
            pass

    _mod("stable_baselines3", PPO=_Dummy)
    _mod("stable_baselines3.common")
    _mod("stable_baselines3.common.utils", configure_logger=None, should_collect_more_steps=None,
         safe_mean=None, obs_as_tensor=None, get_device=None, explained_variance=None,
         get_schedule_fn=None, set_random_seed=None)
    _mod("stable_baselines3.common.policies", ActorCriticPolicy=_Dummy, BasePolicy=_Dummy)
    _mod("stable_baselines3.common.on_policy_algorithm", OnPolicyAlgorithm=_Dummy)
    _mod("stable_baselines3.common.off_policy_algorithm", OffPolicyAlgorithm=_Dummy)
    _mod("stable_baselines3.common.base_class", BaseAlgorithm=_Dummy)
    if "wandb" not in sys.modules:
        try:
            import wandb  # noqa: F401
        except Exception:
            _mod("wandb")

    # bare package so pantheonrl/__init__.py (which pulls unrelated envs) is skipped
    pkg = types.ModuleType("pantheonrl")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "pantheonrl")]
    sys.modules["pantheonrl"] = pkg

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    gc = os.path.join(REFERENCE_ROOT, "gym_cooking")
    if gc not in sys.path:
        sys.path.append(gc)


@contextlib.contextmanager
def _in_cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


@contextlib.contextmanager
def quiet():
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        yield


DEFAULT_AGENT_CFG = {"CAN_MOVE": True, "ALLERGIC": False, "BLIND": False}


def make_namespace(level, num_agents=2, max_num_timesteps=500, communication_on=True,
                   num_communication=10, ego_led=False, fow_radius=2,
                   ego_config=None, partner_config=None):
    """Namespace with the fields the env reads (arglist.py:38-94)."""
    import argparse
    return argparse.Namespace(
        level=level, num_agents=num_agents, max_num_timesteps=max_num_timesteps,
        max_num_subtasks=14, seed=1, model1=None, model2=None, model3=None, model4=None,
        play=False, record=False, with_image_obs=False,
        communication_on=communication_on, num_communication=num_communication,
        ego_led=ego_led, fow_radius=fow_radius,
        ego_config=dict(DEFAULT_AGENT_CFG, **(ego_config or {})),
        partner_config=dict(DEFAULT_AGENT_CFG, **(partner_config or {})),
    )


class LiveReference:
    """Drives the reference env.  2 agents -> the real ``OvercookedMultiEnv``
    wrapper (gym_comm/envs/overcooked_env.py:15-297); 3-4 agents -> the base
    ``OvercookedEnvironment.step`` + ``get_observation2(k)`` directly because the
    wrapper only builds agent-0/agent-1 actions (overcooked_env.py:250-262)."""

    OBS_KEYS = ["agent1_comm", "agent1_location", "agent2_comm", "agent2_location",
                "agent_is_holding", "completed_subtasks", "is_hidden", "object_encodings_x",
                "object_encodings_y", "state_encodings", "timestep"]

    def __init__(self, ns, py_random_seed=None, level_text=None):
        """``level_text``: optional custom level (same 4-phase format).  The reference opens
        ``gym_cooking/utils/levels/<level>.txt`` relative to cwd (overcooked_environment.py:103),
        so a custom level is served from a temp dir used as cwd -- the reference tree is untouched."""
        _install_stubs()
        import random
        self.random = random
        self.ns = ns
        self._level_text = level_text
        if level_text is None:
            self.cwd = REFERENCE_ROOT
        else:
            import tempfile
            self._tmp = tempfile.TemporaryDirectory(prefix="oc_levels_")
            d = os.path.join(self._tmp.name, "gym_cooking", "utils", "levels")
            os.makedirs(d)
            with open(os.path.join(d, ns.level + ".txt"), "w") as f:
                f.write(level_text)
            self.cwd = self._tmp.name
        if py_random_seed is not None:
            random.seed(py_random_seed)
        with _in_cwd(self.cwd), quiet():
            from gym_comm.envs.overcooked_env import OvercookedMultiEnv
            from gym_cooking.utils.world import World
            import gym_cooking.utils.core as Core
            self.Core = Core
            self.NAV = World.NAV_ACTIONS
            self.wrapper = OvercookedMultiEnv(ns)
        self.base = self.wrapper.base_env
        self.n = ns.num_agents

    # -- facts ---------------------------------------------------------
    def subtask_strings(self):
        return [str(s) for s in self.base.all_subtasks]

    def level_text(self):
        if self._level_text is not None:
            return self._level_text
        with open(os.path.join(REFERENCE_ROOT, "gym_cooking/utils/levels", self.ns.level + ".txt")) as f:
            return f.read()

    def object_placements(self):
        """Cells of the dynamic objects in world-dict order (after reset)."""
        out = []
        for objs in self.base.world.objects.values():
            for o in objs:
                if isinstance(o, self.Core.Object):
                    out.append((o.name, tuple(o.location)))
        return out

    # -- stepping ------------------------------------------------------
    def reset(self):
        with _in_cwd(self.cwd), quiet():
            if self.n == 2:
                self.wrapper.multi_reset()
            else:
                self.base.reset()
        return [self.obs(k) for k in range(self.n)]

    def step(self, navs, comms):
        """navs/comms: length-n int lists.  Returns (reward f64, done, sparse int)."""
        with _in_cwd(self.cwd), quiet():
            if self.n == 2:
                _, (r, _r1), done, _ = self.wrapper.multi_step((navs[0], comms[0]), (navs[1], comms[1]))
                return float(r), bool(done)
            # >2 agents: replicate the wrapper's own preamble (overcooked_env.py:227-262)
            import numpy as np
            w = self.wrapper
            C = self.ns.num_communication
            ego = np.zeros(C)
            alt = np.zeros(C)
            if self.ns.communication_on:
                ego[comms[0]] = 1
                if not self.ns.ego_led:
                    alt[comms[1]] = 1
            w.per_agent_communications[0] = ego
            w.per_agent_communications[1] = alt
            ad = {}
            for k in range(self.n):
                cfg = self.ns.ego_config if k == 0 else self.ns.partner_config
                ad["agent-%d" % k] = self.NAV[navs[k]] if cfg["CAN_MOVE"] else (0, 0)
            reward, done, info = self.base.step(ad)
            r = reward - info["agent_0_reward_shaping"] - info["agent_1_reward_shaping"]
            return float(r), bool(done)

    def obs(self, k):
        with quiet():
            return self.wrapper.get_observation2(k, radius=self.ns.fow_radius)

    def flat_obs(self, k):
        """f64 vector in the key-sorted order a gym ``spaces.Dict`` gives."""
        import numpy as np
        o = self.obs(k)
        return np.concatenate([np.asarray(o[key], dtype=np.float64).reshape(-1) for key in self.OBS_KEYS])

    def state_tuple(self):
        """Canonical state for comparison: agents, objects, flags."""
        b = self.base
        agents = [(tuple(a.location), a.holding.full_name if a.holding is not None else None)
                  for a in b.sim_agents]
        objs = []
        for name, lst in b.world.objects.items():
            for o in lst:
                if isinstance(o, self.Core.Object):
                    objs.append((o.full_name, tuple(o.location), bool(o.is_held)))
        return (b.t, agents, objs, list(b.completed_subtasks), list(b.goal_objects_count))
