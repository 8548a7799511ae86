"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/oc_oracle.c (batched C restatement)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "liboc_oracle.so")
ENVS_PER_THREAD = 64
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "oc_oracle.c")
        if not os.path.exists(SO) or os.path.getmtime(src) > os.path.getmtime(SO):
            subprocess.check_call(["make", "-s", "-C", HERE])
        l = C.CDLL(SO)
        l.oco_create.restype = C.c_void_p
        l.oco_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_void_p, C.c_uint64]
        l.oco_destroy.argtypes = [C.c_void_p]
        for name in ("oco_obs_width", "oco_num_random", "oco_state_ints"):
            getattr(l, name).restype = C.c_int
            getattr(l, name).argtypes = [C.c_void_p]
        l.oco_threads.restype = C.c_int
        l.oco_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.oco_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        l.oco_rollout.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.oco_throughput.restype = C.c_longlong
        l.oco_throughput.argtypes = [C.c_void_p, C.c_double]
        l.oco_get_state.argtypes = [C.c_void_p, C.c_void_p]
        l.oco_set_clocks.argtypes = [C.c_void_p, C.c_void_p]
        _lib = l
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class COracle:
    """N independent envs.  Same constructor vocabulary as SpecEnv."""

    def __init__(self, level_text, subtasks, num_envs, num_agents=2, max_num_timesteps=500,
                 communication_on=True, num_communication=10, ego_led=False, fow_radius=2,
                 ego_config=None, partner_config=None, seed=0):
        d = {"CAN_MOVE": True, "ALLERGIC": False, "BLIND": False}
        cfgs = [dict(d, **(ego_config or {})), dict(d, **(partner_config or {}))]
        flags = np.array([[int(cfgs[min(k, 1)][f]) for f in ("CAN_MOVE", "ALLERGIC", "BLIND")]
                          for k in range(num_agents)], dtype=np.int32)
        self.l = lib()
        self.h = self.l.oco_create(level_text.encode(), ";".join(subtasks).encode(), num_envs, num_agents,
                                   max_num_timesteps, int(communication_on), num_communication, int(ego_led),
                                   fow_radius, _p(flags), seed)
        if not self.h:
            raise ValueError("oc_oracle could not parse the level / subtasks")
        self.N, self.A = num_envs, num_agents
        self.F = self.l.oco_obs_width(self.h)
        self.R = self.l.oco_num_random(self.h)
        self.S = len(subtasks)
        self.obs = np.zeros((num_envs, num_agents, self.F), dtype=np.float64)
        self.reward = np.zeros(num_envs, dtype=np.float64)
        self.done = np.zeros(num_envs, dtype=np.uint8)

    def reset(self, mask=None, placements=None):
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
        if placements is not None:
            placements = np.ascontiguousarray(placements, dtype=np.int32).reshape(self.N, self.R)
        self.l.oco_reset(self.h, _p(mask), _p(placements), _p(self.obs))
        return self.obs

    def step(self, actions, auto_reset=False, term_obs=None):
        actions = np.ascontiguousarray(actions, dtype=np.int32).reshape(self.N, self.A, 2)
        self.l.oco_step(self.h, _p(actions), _p(self.obs), _p(self.reward), _p(self.done), int(auto_reset), _p(term_obs))
        return self.obs, self.reward, self.done

    def rollout(self, n_steps, want_obs=True, want_actions=False):
        obs = np.zeros((n_steps, self.N, self.A, self.F), dtype=np.float64) if want_obs else None
        rew = np.zeros((n_steps, self.N), dtype=np.float64)
        done = np.zeros((n_steps, self.N), dtype=np.uint8)
        acts = np.zeros((n_steps, self.N, self.A, 2), dtype=np.int32) if want_actions else None
        self.l.oco_rollout(self.h, n_steps, _p(obs), _p(rew), _p(done), _p(acts))
        return obs, rew, done, acts

    def set_clocks(self, t):
        """Episode clock of every env := t[e] (what `stagger_clocks` does to the device state)."""
        t = np.ascontiguousarray(t, dtype=np.uint32).reshape(self.N)
        self.l.oco_set_clocks(self.h, _p(t))

    def state(self):
        n = self.l.oco_state_ints(self.h)
        out = np.zeros((self.N, n), dtype=np.int32)
        self.l.oco_get_state(self.h, _p(out))
        A, S = self.A, self.S
        p = 2 + 2 * A
        return dict(t=out[:, 0], episodes=out[:, 1], agents=out[:, 2:p].reshape(self.N, A, 2),
                    completed=out[:, p:p + S], counts=out[:, p + S:p + 2 * S],
                    comm=out[:, p + 2 * S:p + 2 * S + 2], last_completed=out[:, p + 2 * S + 2],
                    objs=out[:, p + 2 * S + 3:].reshape(self.N, -1, 5))

    def close(self):
        if self.h:
            self.l.oco_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def throughput(ns, seconds):
    """(env_steps, seconds, threads) of random-action stepping incl. obs + resets over all host cores."""
    import time
    from gym_comm_b200 import levels_data   # data only
    text = levels_data.LEVELS[ns.level]
    subtasks = levels_data.SUBTASKS[tuple(text.split("\n\n")[1].split("\n"))]
    threads = lib().oco_threads()
    env = COracle(text, subtasks, threads * ENVS_PER_THREAD, num_agents=ns.num_agents,
                  max_num_timesteps=ns.max_num_timesteps, communication_on=ns.communication_on,
                  num_communication=ns.num_communication, ego_led=ns.ego_led, fow_radius=ns.fow_radius,
                  ego_config=ns.ego_config, partner_config=ns.partner_config, seed=7)
    lib().oco_throughput(env.h, 0.5)   # warm-up
    t0 = time.perf_counter()
    steps = lib().oco_throughput(env.h, float(seconds))
    dt = time.perf_counter() - t0
    env.close()
    return int(steps), dt, threads
