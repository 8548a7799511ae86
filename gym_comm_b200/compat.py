"""Registration shim for a box that HAS gym + PantheonRL (level 1): the drop-in behind the
reference's own `SimultaneousEnv` base class, so that PantheonRL's `MultiAgentEnv.step` / `reset` / partner
handling (pantheonrl/common/multiagentenv.py:38-243) run unchanged on top of the CUDA env.

    # gym_comm/__init__.py
    register(id='OvercookedMultiCommEnv-v0', entry_point=gym_comm_b200.compat.gym_env_class())

Nothing here is imported by the rest of the package: `pantheonrl` is looked up only when `gym_env_class()` is
called (this image has neither gym nor pantheonrl; the build container tests it against the reference's own
`pantheonrl` package, tests/test_compat_pantheonrl.py).
"""
from __future__ import annotations


def gym_env_class(b200_env_cls=None):
    """-> a `SimultaneousEnv` subclass with the constructor of `gym_comm.envs.OvercookedMultiEnv`
    (gym_comm/envs/overcooked_env.py:16-18) whose `multi_step` / `multi_reset` go to the B200 env
    (`b200_env_cls`: the class to wrap, default `gym_comm_b200.OvercookedMultiEnv`)."""
    from pantheonrl.common.multiagentenv import SimultaneousEnv       # the reference's own base class

    if b200_env_cls is None:
        from .vec_env import OvercookedMultiEnv as b200_env_cls
    _B200Env = b200_env_cls

    class GymOvercookedMultiEnv(SimultaneousEnv):
        def __init__(self, arglist, ego_agent_idx: int = 0, baselines: bool = False, **backend):
            super().__init__()
            self._env = _B200Env(arglist, ego_agent_idx=ego_agent_idx, **backend)     # CUDA handle, one env
            self.arglist = self._env.arglist
            self.ego_agent_idx = ego_agent_idx
            self.observation_space = self._env.observation_space
            self.action_space = self._env.action_space
            self.lA = self._env.lA

        def multi_step(self, ego_action, alt_action):
            return self._env.multi_step(ego_action, alt_action)

        def multi_reset(self):
            return self._env.multi_reset()

        def get_observation2(self, agent_idx, radius=None):
            return self._env.get_observation2(agent_idx, radius)

        def render(self, mode="human", close=False):
            return self._env.render(mode, close)

        def close(self):
            self._env.close()

    return GymOvercookedMultiEnv
