"""Batched Overcooked env on the GPU behind the reference's env API.

``OvercookedVecEnv``  -- E lock-step envs, state in HBM, every call one CUDA launch through the
                         C ABI (include/overcooked_b200.h); tensors in, tensors out, no host sync.
``OvercookedMultiEnv`` -- the reference's 2-player ``SimultaneousEnv`` surface
                         (gym_comm/envs/overcooked_env.py:15-297: ``multi_step`` / ``multi_reset`` /
                         ``get_observation2`` with the 11-key dict observation) on top of it.

The dynamics live in gym_comm_b200/csrc (hand-written sm_100a CUDA).  Nothing here computes
env logic on the host, and nothing falls back to a CPU implementation.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _cabi
from .arglist import normalize
from .level_compiler import NAV_ACTIONS, CompiledLevel, compile_level


from .spaces import Box, Dict, MultiBinary, MultiDiscrete, make_spaces  # noqa: F401  (re-exported)


class OvercookedVecEnv:
    """E Overcooked envs stepped in lock-step on one GPU.

    ``arglist``: Namespace / dict with the reference's keys (arglist.py:96-121).
    Observations are flat float32 rows ``[E, A, F]`` (F = 23 + S + 2C) in the key-sorted order of
    the reference's Dict space -- what ``FlattenedDictExtractor`` would concatenate
    (gym_comm/extractors/CustomExtractor.py:119-128); ``obs_dict`` gives zero-copy per-key views.
    """

    def __init__(self, arglist, num_envs: int = 1, device="cuda", seed: int = 0, auto_reset: bool = True,
                 level_text: Optional[str] = None, subtasks=None, lib: Optional[_cabi.OcLibrary] = None):
        self.arglist = normalize(arglist)
        a = self.arglist
        self.lib = lib if lib is not None else _cabi.default_library()
        self.device = torch.device(device)
        self._require_backend()
        self.num_envs = int(num_envs)
        self.num_agents = int(a.num_agents)
        self.auto_reset = bool(auto_reset)
        self.level: CompiledLevel = compile_level(a.level, self.num_agents, level_text=level_text, subtasks=subtasks)
        cfg, self._keep = _cabi.make_config(
            self.level, num_envs=self.num_envs, num_agents=self.num_agents,
            max_num_timesteps=a.max_num_timesteps, num_communication=a.num_communication,
            communication_on=a.communication_on, ego_led=a.ego_led, fow_radius=a.fow_radius,
            ego_config=a.ego_config, partner_config=a.partner_config, seed=seed)
        self._handle = C.c_void_p()
        with self._device_guard():
            self.lib.check(self.lib.create(C.byref(cfg), C.byref(self._handle)), "oc_create")
        self.obs_width = self.lib.obs_width(self._handle)
        off = (C.c_int32 * _cabi.OC_NUM_OBS_KEYS)()
        size = (C.c_int32 * _cabi.OC_NUM_OBS_KEYS)()
        self.lib.check(self.lib.obs_layout(self._handle, off, size), "oc_obs_layout")
        self.obs_layout = {k: slice(off[i], off[i] + size[i]) for i, k in enumerate(_cabi.OBS_KEYS)}
        self.observation_space, self.action_space = make_spaces(self.level, a.num_communication)
        E, A, F = self.num_envs, self.num_agents, self.obs_width
        kw = dict(device=self.device)
        self.obs = torch.zeros((E, A, F), dtype=torch.float32, **kw)
        self.rewards = torch.zeros((E, A), dtype=torch.float32, **kw)
        self.rewards64 = torch.zeros((E,), dtype=torch.float64, **kw)
        self.dones = torch.zeros((E,), dtype=torch.uint8, **kw)
        self._closed = False

    # ------------------------------------------------------------------ plumbing
    def _require_backend(self):
        """The only backend is the CUDA library on a CUDA device."""
        if not isinstance(self.lib, _cabi.OcLibrary):
            raise RuntimeError("lib must be a gym_comm_b200._cabi.OcLibrary (liboc_b200.so); there is no other backend")
        if self.device.type != "cuda":
            raise RuntimeError("OvercookedVecEnv runs on a CUDA device only (no CPU fallback)")
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device available; the Overcooked kernels are sm_100a CUDA only")

    def _device_guard(self):
        return torch.cuda.device(self.device)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _ptr(t: Optional[torch.Tensor]):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _check_tensor(self, t, shape, dtype, name):
        if t.device != self.obs.device or t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
            raise ValueError("%s must be a contiguous %s tensor of shape %s on %s (got %s %s on %s)" %
                             (name, dtype, tuple(shape), self.obs.device, t.dtype, tuple(t.shape), t.device))

    # ------------------------------------------------------------------ env API
    def reset(self, mask: Optional[torch.Tensor] = None, placements: Optional[torch.Tensor] = None,
              obs_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Reset all envs (or those with ``mask[e] != 0``).  ``placements`` int32 [E, R] pins the cells
        of the R random (phase-4) objects; default: drawn on the device."""
        obs = self.obs if obs_out is None else obs_out
        self._check_tensor(obs, self.obs.shape, torch.float32, "obs_out")
        if mask is not None:
            self._check_tensor(mask, (self.num_envs,), torch.uint8, "mask")
        if placements is not None:
            self._check_tensor(placements, (self.num_envs, self.level.num_random), torch.int32, "placements")
        with self._device_guard():
            self.lib.check(self.lib.reset(self._handle, self._ptr(mask), self._ptr(placements), self._ptr(obs),
                                          self._stream()), "oc_reset")
        return obs

    _CHAIN = {None: 0, "head": _cabi.OC_FLAG_CHAIN_HEAD, "next": _cabi.OC_FLAG_CHAINED}

    def step(self, actions: torch.Tensor, obs_out=None, rew_out=None, done_out=None, term_obs_out=None,
             want_f64: bool = False, chain=None):
        """``actions`` int32 [E, A, 2] = (nav in [0,4), comm in [0,C)) per agent.  Returns
        (obs [E,A,F] f32, rewards [E,A] f32, dones [E] u8); with ``want_f64`` the reward in the
        reference's own f64 is also left in ``self.rewards64``.  Asynchronous on the current
        stream; outputs may be caller-provided rollout-buffer slots.
        ``chain``: ``"head"`` / ``"next"`` for a run of steps whose actions are all in memory beforehand
        (OC_FLAG_CHAIN_HEAD / OC_FLAG_CHAINED, include/overcooked_b200.h): the steps then overlap, each
        starting as soon as the previous one has stored its states; consecutive steps must write
        different output tensors."""
        obs = self.obs if obs_out is None else obs_out
        rew = self.rewards if rew_out is None else rew_out
        done = self.dones if done_out is None else done_out
        self._check_tensor(actions, (self.num_envs, self.num_agents, 2), torch.int32, "actions")
        self._check_tensor(obs, self.obs.shape, torch.float32, "obs_out")
        self._check_tensor(rew, self.rewards.shape, torch.float32, "rew_out")
        self._check_tensor(done, self.dones.shape, torch.uint8, "done_out")
        if term_obs_out is not None:
            self._check_tensor(term_obs_out, self.obs.shape, torch.float32, "term_obs_out")
        flags = (_cabi.OC_FLAG_AUTO_RESET if self.auto_reset else 0) | self._CHAIN[chain]
        with self._device_guard():
            self.lib.check(self.lib.step(self._handle, self._ptr(actions), self._ptr(obs), self._ptr(rew),
                                         self._ptr(self.rewards64) if want_f64 else None, self._ptr(done),
                                         self._ptr(term_obs_out), flags, self._stream()), "oc_step")
        return obs, rew, done

    def compact_buffers(self):
        """Output tensors of `step_i8` / `reset_i8`: (obs int8 [E, A, F-1], timestep f32 [E])."""
        E, A, F = self.num_envs, self.num_agents, self.obs_width
        return (torch.zeros((E, A, F - 1), dtype=torch.int8, device=self.device),
                torch.zeros((E,), dtype=torch.float32, device=self.device))

    def reset_i8(self, obs_out: torch.Tensor, timestep_out: torch.Tensor, mask=None, placements=None):
        """`reset` with the observations in the compact integer format (`oc_reset_i8`): the reset kernel writes
        int8 [E, A, F-1] rows and the f32 [E] clock itself."""
        E, A, F = self.num_envs, self.num_agents, self.obs_width
        self._check_tensor(obs_out, (E, A, F - 1), torch.int8, "obs_out")
        self._check_tensor(timestep_out, (E,), torch.float32, "timestep_out")
        if mask is not None:
            self._check_tensor(mask, (E,), torch.uint8, "mask")
        if placements is not None:
            self._check_tensor(placements, (E, self.level.num_random), torch.int32, "placements")
        with self._device_guard():
            self.lib.check(self.lib.reset_i8(self._handle, self._ptr(mask), self._ptr(placements), self._ptr(obs_out),
                                             self._ptr(timestep_out), self._stream()), "oc_reset_i8")
        return obs_out, timestep_out

    def step_i8(self, actions: torch.Tensor, obs_out: torch.Tensor, timestep_out: torch.Tensor, rew_out=None,
                done_out=None, term_obs_out=None, term_timestep_out=None, want_f64: bool = False, chain=None):
        """`step` with the observations in the compact integer format (`oc_step_i8`): int8 [E, A, F-1] rows +
        f32 [E] clock written by the step kernel itself (a quarter of the bytes of the float rows).
        ``actions``: int32 [E, A, 2] or uint8 [E, A, 2]; ``rew_out``: f32 [E, A] (default) or f32 [E]."""
        E, A, F = self.num_envs, self.num_agents, self.obs_width
        rew = self.rewards if rew_out is None else rew_out
        done = self.dones if done_out is None else done_out
        flags = (_cabi.OC_FLAG_AUTO_RESET if self.auto_reset else 0) | self._CHAIN[chain]
        if actions.dtype == torch.uint8:
            self._check_tensor(actions, (E, A, 2), torch.uint8, "actions")
            flags |= _cabi.OC_FLAG_ACTIONS_U8
        else:
            self._check_tensor(actions, (E, A, 2), torch.int32, "actions")
        self._check_tensor(obs_out, (E, A, F - 1), torch.int8, "obs_out")
        self._check_tensor(timestep_out, (E,), torch.float32, "timestep_out")
        if rew.dim() == 1:
            self._check_tensor(rew, (E,), torch.float32, "rew_out")
            flags |= _cabi.OC_FLAG_REWARD_PER_ENV
        else:
            self._check_tensor(rew, (E, A), torch.float32, "rew_out")
        self._check_tensor(done, (E,), torch.uint8, "done_out")
        if term_obs_out is not None:
            self._check_tensor(term_obs_out, (E, A, F - 1), torch.int8, "term_obs_out")
        if term_timestep_out is not None:
            self._check_tensor(term_timestep_out, (E,), torch.float32, "term_timestep_out")
        with self._device_guard():
            self.lib.check(self.lib.step_i8(self._handle, self._ptr(actions), self._ptr(obs_out), self._ptr(timestep_out),
                                            self._ptr(rew), self._ptr(self.rewards64) if want_f64 else None,
                                            self._ptr(done), self._ptr(term_obs_out), self._ptr(term_timestep_out),
                                            flags, self._stream()), "oc_step_i8")
        return obs_out, timestep_out, rew, done

    def rollout(self, n_steps: int, obs_out=None, rew_out=None, done_out=None, actions_out=None):
        """Fused synthetic rollout: ``n_steps`` steps in ONE launch, uniform random actions from
        Philox on the device, auto-reset on.  Outputs are [n_steps, E, ...] (any may be None)."""
        E, A, F = self.num_envs, self.num_agents, self.obs_width
        if obs_out is not None:
            self._check_tensor(obs_out, (n_steps, E, A, F), torch.float32, "obs_out")
        if rew_out is not None:
            self._check_tensor(rew_out, (n_steps, E, A), torch.float32, "rew_out")
        if done_out is not None:
            self._check_tensor(done_out, (n_steps, E), torch.uint8, "done_out")
        if actions_out is not None:
            self._check_tensor(actions_out, (n_steps, E, A, 2), torch.int32, "actions_out")
        with self._device_guard():
            self.lib.check(self.lib.rollout(self._handle, int(n_steps), self._ptr(obs_out), self._ptr(rew_out),
                                            self._ptr(done_out), self._ptr(actions_out), self._stream()), "oc_rollout")

    def replay(self, actions: torch.Tensor, obs_out=None, rew_out=None, done_out=None):
        """Open-loop replay of an action sequence int32 [n_steps, E, A, 2] in ONE launch (auto-reset
        on); outputs as for `rollout`."""
        n_steps = int(actions.shape[0])
        E, A, F = self.num_envs, self.num_agents, self.obs_width
        self._check_tensor(actions, (n_steps, E, A, 2), torch.int32, "actions")
        if obs_out is not None:
            self._check_tensor(obs_out, (n_steps, E, A, F), torch.float32, "obs_out")
        if rew_out is not None:
            self._check_tensor(rew_out, (n_steps, E, A), torch.float32, "rew_out")
        if done_out is not None:
            self._check_tensor(done_out, (n_steps, E), torch.uint8, "done_out")
        with self._device_guard():
            self.lib.check(self.lib.replay(self._handle, n_steps, self._ptr(actions), self._ptr(obs_out), self._ptr(rew_out),
                                           self._ptr(done_out), self._stream()), "oc_replay")

    # ------------------------------------------------------------------ state / stats
    def get_state(self) -> torch.Tensor:
        st = torch.zeros((self.num_envs, _cabi.OC_STATE_WORDS), dtype=torch.int32, device=self.device)
        with self._device_guard():
            self.lib.check(self.lib.get_state(self._handle, self._ptr(st), self._stream()), "oc_get_state")
        return st

    def set_state(self, st: torch.Tensor):
        self._check_tensor(st, (self.num_envs, _cabi.OC_STATE_WORDS), torch.int32, "state")
        with self._device_guard():
            self.lib.check(self.lib.set_state(self._handle, self._ptr(st), self._stream()), "oc_set_state")

    def stats(self):
        ep = torch.zeros((self.num_envs,), dtype=torch.int32, device=self.device)
        lc = torch.zeros((self.num_envs,), dtype=torch.int32, device=self.device)
        with self._device_guard():
            self.lib.check(self.lib.get_stats(self._handle, self._ptr(ep), self._ptr(lc), self._stream()), "oc_get_stats")
        return {"episodes": ep, "num_completed_subtasks": lc}

    def launch_count(self) -> int:
        return int(self.lib.launch_count(self._handle))

    def obs_dict(self, obs: Optional[torch.Tensor] = None) -> dict:
        """Zero-copy per-key views ``[..., size]`` of flat observation rows."""
        obs = self.obs if obs is None else obs
        return {k: obs[..., s] for k, s in self.obs_layout.items()}

    def pack_obs_i8(self, obs: Optional[torch.Tensor] = None, out=None, timestep_out=None):
        """Float rows [E, A, F] -> the compact integer format of `oc_pack_obs_i8` (include/overcooked_b200.h):
        ``(int8 [E, A, F-1], f32 [E] clock)`` on the same device, e.g. before shipping a rollout to a host."""
        obs = self.obs if obs is None else obs
        E, A, F = self.num_envs, self.num_agents, self.obs_width
        self._check_tensor(obs, (E, A, F), torch.float32, "obs")
        out = torch.empty((E, A, F - 1), dtype=torch.int8, device=self.device) if out is None else out
        ts = torch.empty((E,), dtype=torch.float32, device=self.device) if timestep_out is None else timestep_out
        self._check_tensor(out, (E, A, F - 1), torch.int8, "out")
        self._check_tensor(ts, (E,), torch.float32, "timestep_out")
        with self._device_guard():
            self.lib.check(self.lib.pack_obs_i8(self._handle, self._ptr(obs), self._ptr(out), self._ptr(ts), self._stream()),
                           "oc_pack_obs_i8")
        return out, ts

    def decode_state(self, st: Optional[torch.Tensor] = None):
        """Packed state -> readable numpy fields (layout: gym_comm_b200/csrc/oc_params.h)."""
        w = (self.get_state() if st is None else st).cpu().numpy().view(np.uint32)
        A, W = self.num_agents, self.level.width
        cells = np.stack([(w[:, 4] >> (8 * k)) & 0xFF for k in range(A)], 1)
        obj = w[:, 8:14]
        out = dict(
            t=w[:, 0] & 0xFFFF, next_stamp=(w[:, 0] >> 16) & 0xFF, nkeys=w[:, 0] >> 24,
            episodes=w[:, 1], completed=w[:, 2], countbits=w[:, 3],
            agent_cell=cells, agent_x=cells % W, agent_y=cells // W,
            last_completed=w[:, 5] & 0xFF,
            ranks=w[:, 6].astype(np.uint64) | (w[:, 7].astype(np.uint64) << np.uint64(32)),
            obj_contents=obj & 0xF, obj_chopped=(obj >> 4) & 7, obj_holder=(obj >> 8) & 7,
            obj_cell=(obj >> 16) & 0xFF, obj_stamp=obj >> 24,
            comm0=w[:, 14] & 0xFFFF, comm1=w[:, 14] >> 16)
        return out

    def render(self, index: int = 0, state=None) -> str:
        """ASCII picture of env `index` in the reference's `str(OvercookedEnvironment)` format
        (overcooked_environment.py:62-65, world.py:36-46, core.py:277-279): one character cell + a
        space per tile; objects print as `1t` (fresh) / `2t` (chopped) / `p`, merged ones joined
        by '-' in name order; agents print their index and hide what is under them."""
        d = self.decode_state(state)
        lv = self.level
        tile_rep = {0: " ", 1: "-", 2: "/", 3: "*"}
        grid = [[tile_rep[int(lv.tiles[lv.cell(x, y)])] for x in range(lv.width)] for y in range(lv.height)]
        names = [(2, "l"), (4, "o"), (8, "p"), (1, "t")]                 # alphabetical: Lettuce Onion Plate Tomato
        ranks = int(d["ranks"][index])
        objs = []
        for s in range(6):
            c = int(d["obj_contents"][index, s])
            if c:
                ch = int(d["obj_chopped"][index, s])
                txt = "-".join(("p" if b == 8 else "%d%s" % (2 if ch & b else 1, r)) for b, r in names if c & b)
                objs.append((((ranks >> (4 * c)) & 15, int(d["obj_stamp"][index, s])), int(d["obj_cell"][index, s]), txt, c))
        objs.sort(key=lambda o: o[0])                                    # world.objects iteration order
        for _, cell, txt, c in objs:
            grid[cell // lv.width][cell % lv.width] = txt
        for _, cell, txt, c in objs:                                     # Tomato objects are drawn again last (world.py:44-45)
            if c == 1:
                grid[cell // lv.width][cell % lv.width] = txt
        for k in range(self.num_agents):
            grid[int(d["agent_y"][index, k])][int(d["agent_x"][index, k])] = str(k)
        return "\n".join("".join(c + " " for c in row) for row in grid)

    def close(self):
        if not self._closed and self._handle:
            self.lib.destroy(self._handle)
            self._closed = True

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OvercookedMultiEnv:
    """The reference's ``OvercookedMultiEnv`` surface (gym_comm/envs/overcooked_env.py:15-297).

    ``multi_step(ego_action, alt_action)`` takes ``(nav, comm)`` per player and returns
    ``((obs0, obs1), (r, r), done, {})`` with the reference's 11-key numpy dict observations and the
    reward as a Python float carrying the reference's f64 value; ``multi_reset()`` returns
    ``(obs0, obs1)``.  It is a thin view on a 1-env (or env ``index`` of an E-env)
    ``OvercookedVecEnv``; episodes do NOT auto-reset here (the reference leaves that to the caller,
    pantheonrl/common/multiagentenv.py:217-243)."""

    n_players = 2
    _vec_cls = OvercookedVecEnv

    def __init__(self, arglist, ego_agent_idx: int = 0, device="cuda", seed: int = 0,
                 level_text: Optional[str] = None, subtasks=None, lib=None):
        self.arglist = normalize(arglist)
        if self.arglist.num_agents != 2:
            raise ValueError("OvercookedMultiEnv is a 2-player SimultaneousEnv (multiagentenv.py:390-393); "
                             "use OvercookedVecEnv for 3-4 agents")
        if ego_agent_idx != 0:
            raise ValueError("only ego_agent_idx=0 is supported (the only value the reference's trainer uses)")
        self.ego_agent_idx = ego_agent_idx
        self.vec = self._vec_cls(self.arglist, num_envs=1, device=device, seed=seed, auto_reset=False,
                                    level_text=level_text, subtasks=subtasks, lib=lib)
        self.observation_space = self.vec.observation_space
        self.action_space = self.vec.action_space
        self.lA = len(NAV_ACTIONS)
        self._actions = torch.zeros((1, 2, 2), dtype=torch.int32, device=self.vec.device)
        self.multi_reset()

    def _split(self, obs_row: np.ndarray):
        out = {}
        t = int(self.vec.decode_state()["t"][0])
        for k, s in self.vec.obs_layout.items():
            v = obs_row[s]
            if k == "timestep":       # the reference returns the f64 quotient (overcooked_env.py:146), not its f32 rounding
                out[k] = np.array((t / self.arglist.max_num_timesteps,))
                continue
            if k in ("object_encodings_x", "object_encodings_y", "state_encodings", "is_hidden",
                     "completed_subtasks"):
                v = v.astype(np.int64)
            elif k == "agent_is_holding":
                v = v.astype(bool) if not self.arglist.ego_config["BLIND"] else v.astype(np.int64)
            elif k in ("agent1_location", "agent2_location"):
                v = v.astype(np.int64)
            else:
                v = v.astype(np.float64)
            out[k] = v
        return out

    def get_observation2(self, agent_idx, radius=None):
        """Observation of ``agent_idx`` at the configured ``fow_radius`` (the value training sees,
        overcooked_env.py:282,297)."""
        return self._split(self.vec.obs[0, agent_idx].cpu().numpy())

    def multi_step(self, ego_action, alt_action):
        acts = np.array([[[int(ego_action[0]), int(ego_action[1])], [int(alt_action[0]), int(alt_action[1])]]],
                        dtype=np.int32)
        self._actions.copy_(torch.from_numpy(acts))
        obs, _, done = self.vec.step(self._actions, want_f64=True)
        r = float(self.vec.rewards64[0].item())
        o = obs[0].cpu().numpy()
        return (self._split(o[0]), self._split(o[1])), (r, r), bool(done[0].item()), {}

    def multi_reset(self, placements=None):
        pl = None
        if placements is not None:
            pl = torch.tensor(np.asarray(placements, dtype=np.int32).reshape(1, -1), device=self.vec.device)
        o = self.vec.reset(placements=pl)[0].cpu().numpy()
        return (self._split(o[0]), self._split(o[1]))

    # ---- MultiAgentEnv surface (pantheonrl/common/multiagentenv.py:72-243), single env, ego index 0
    class Observation:
        """pantheonrl.common.observation.Observation: what a partner's `get_action` receives."""

        def __init__(self, obs):
            self.obs, self.state, self.action_mask = obs, obs, None

    def add_partner_agent(self, agent, player_num: int = 1) -> None:
        if player_num != 1:
            raise ValueError("Ego agent is not set by the environment")      # PlayerException (:86-87)
        self.__dict__.setdefault("partners", []).append(agent)
        self.__dict__.setdefault("partnerid", 0)

    def getDummyEnv(self, player_num: int):
        return self

    # partner selection (multiagentenv.py:103-147); one partner slot, as in every 2-player PantheonRL env
    def set_partnerid(self, agent_id: int, player_num: int = 1) -> None:
        if player_num != 1:
            raise ValueError("Ego agent is not set by the environment")
        partners = self.__dict__.get("partners", [])
        assert 0 <= agent_id < len(partners)
        self.partnerid = int(agent_id)

    def resample_random(self) -> None:
        self.partnerid = int(np.random.randint(len(self.partners)))

    def resample_round_robin(self) -> None:
        self.partnerid = (self.partnerid + 1) % len(self.partners)

    def set_resample_policy(self, resample_policy: str) -> None:
        """"default" / "robin" (round robin, the 2-player default) or "random"."""
        if resample_policy in ("default", "robin"):
            self.resample_partner = self.resample_round_robin
        elif resample_policy == "random":
            self.resample_partner = self.resample_random
        else:
            raise ValueError("Invalid resampling policy: %s" % resample_policy)      # PlayerException (:145-147)

    # SimultaneousEnv.n_step / n_reset (multiagentenv.py:395-409): both players move on every step
    def n_step(self, actions):
        (obs0, obs1), r, d, i = self.multi_step(actions[0], actions[1])
        return (0, 1), (self.Observation(obs0), self.Observation(obs1)), r, d, i

    def n_reset(self):
        obs0, obs1 = self.multi_reset()
        return (0, 1), (self.Observation(obs0), self.Observation(obs1))

    def cost_fn(self):
        return 1                                                                      # overcooked_env.py:204-205

    def render(self, mode="human", close=False):
        """Prints the board like `print(str(self.base_env))` (overcooked_env.py:299-300; the reference then calls a
        `get_observation` that does not exist and raises) followed by the ego observation."""
        print(self.vec.render(0))
        print(self.get_observation2(self.ego_agent_idx))

    def set_ego_extractor(self, ego_extractor) -> None:
        self.ego_extractor = ego_extractor

    def reset(self, placements=None):
        """MultiAgentEnv.reset (:217-243): round-robin partner resampling, first ego observation."""
        partners = self.__dict__.get("partners", [])
        if partners:
            self.__dict__.get("resample_partner", self.resample_round_robin)()   # (:228-229), round robin by default
        self._obs = self.multi_reset(placements)
        self._should_update, self._total_rews = False, [0.0, 0.0]
        self._old_ego_obs = self._obs[0]
        return self.__dict__.get("ego_extractor", lambda o: o)(self._obs[0])

    def step(self, action):
        """MultiAgentEnv.step (:172-215): the partner acts inside the step on the observation it saw
        last, gets `update(reward, done)`; on done the PREVIOUS ego observation is returned."""
        partners = self.__dict__.get("partners", [])
        if not partners:
            raise RuntimeError("add_partner_agent first")
        if "_obs" not in self.__dict__:
            self.reset()
        agent = partners[self.partnerid]
        alt = agent.get_action(self.Observation(self._obs[1]))
        if not self._should_update:
            agent.update(self._total_rews[1], False)                             # _get_actions (:157-159)
        self._should_update = True
        obs, rews, done, info = self.multi_step(action, alt)
        info["_partnerid"] = [self.partnerid]
        agent.update(rews[1], done)                                              # _update_players (:163-170)
        self._total_rews = [self._total_rews[0] + rews[0], self._total_rews[1] + rews[1]]
        extract = self.__dict__.get("ego_extractor", lambda o: o)
        if done:
            return extract(self._old_ego_obs), rews[0], done, info               # (:204-208)
        self._obs = obs
        self._old_ego_obs = obs[0]
        return extract(obs[0]), rews[0], done, info

    def close(self):
        self.vec.close()
