"""The batched Overcooked env for a caller whose buffers live in HOST memory -- numpy in, numpy out,
no torch anywhere: the C ABI's host-buffer entry points (`oc_reset_host` / `oc_step_host`,
include/overcooked_b200.h) behind the SB3 `VecEnv` calling convention.

This is what stands where the reference's `DummyVecEnv([lambda: OvercookedMultiEnv])` stands when
the learner stays on the CPU (sb3_contrib/ppo_recurrent/ppo_recurrent.py:233-252 reads numpy
observations and writes numpy actions): every `step` copies the actions host->device, runs the CUDA
step kernel, copies observations / rewards / dones device->host and synchronises, so it is
PCIe-bound (`bench.py` reports it as `e2e`).  A learner on the GPU should use
`gym_comm_b200.vec_env.OvercookedVecEnv` and never leave the device.

The compute is `liboc_b200.so` (sm_100a CUDA); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _cabi
from .arglist import normalize
from .level_compiler import CompiledLevel, compile_level
from .spaces import make_spaces


def finished_indices(done_padded: np.ndarray, num_envs: int) -> list:
    """Indices of the set flags in a uint8 done buffer whose length is padded to a multiple of 8.
    Finished envs are rare in a large batch (about E / T per step), so the flags are scanned eight at a
    time as uint64 words and only the non-zero words are opened up -- several times cheaper than
    `np.flatnonzero` over the bytes, which is what a 65 536-env step would otherwise spend its Python time on."""
    w = np.flatnonzero(done_padded.view(np.uint64))
    if w.size == 0:
        return []
    if w.size * 16 > done_padded.size:                     # dense (lock-step envs hitting the time limit together)
        return np.flatnonzero(done_padded[:num_envs]).tolist()
    r, c = np.nonzero(done_padded.reshape(-1, 8)[w])
    return (w[r] * 8 + c).tolist()


class OvercookedHostVecEnv:
    """E lock-step envs on one GPU, host (numpy) buffers.

    ``reset() -> obs [E, A, F] f32``;
    ``step(actions [E, A, 2]) -> (obs [E, A, F] f32, rewards [E, A] f32, dones [E] bool, infos)``
    with SB3's auto-reset contract: for an env that finished, ``obs`` is the first observation of its
    next episode and ``infos[e]["terminal_observation"]`` the last one of the finished episode (a view of
    ``env.terminal_obs[e]``; with very large batches read ``env.terminal_obs`` and ``done`` directly instead
    of walking the list).
    ``obs_format="i8"`` switches to the compact integer format and the ONE-BLOCK host path
    (`oc_step_host_block`, include/overcooked_b200.h): ``obs`` is int8 [E, A, F-1] (every key but the clock --
    the reference builds these keys as integer arrays, overcooked_env.py:145-157), the clock is
    ``env.timestep`` f32 [E] (after a finished episode: ``infos[e]["terminal_timestep"]``), the reward is one
    f32 per env (``rewards`` is its read-only [E, A] broadcast view; the reference hands both players the same
    number, overcooked_env.py:282) and the actions cross PCIe as two bytes per agent.  Everything a step returns
    sits in one page-locked block filled by one device->host copy; the step kernel writes the compact rows
    itself (no float rows, no repack).
    ``obs_float()`` rebuilds the float rows, ``obs_dict()`` gives the reference's per-key view in both formats.
    The returned arrays (and the `infos` list) are the env's own pinned buffers, overwritten by the next
    call (copy them to keep them, as SB3's rollout buffer does) and invalid after `close()`.
    """

    def __init__(self, arglist, num_envs: int = 1, device_index: int = 0, seed: int = 0, auto_reset: bool = True,
                 terminal_observations: bool = True, level_text: Optional[str] = None, subtasks=None,
                 lib: Optional[_cabi.OcLibrary] = None, obs_format: str = "f32"):
        self.arglist = normalize(arglist)
        a = self.arglist
        self.lib = lib if lib is not None else _cabi.default_library()
        self._require_backend()
        if obs_format not in ("f32", "i8"):
            raise ValueError("obs_format must be 'f32' or 'i8'")
        self.obs_format = obs_format
        self.device_index = int(device_index)
        self.num_envs, self.num_agents = int(num_envs), int(a.num_agents)
        self.auto_reset = bool(auto_reset)
        self.level: CompiledLevel = compile_level(a.level, self.num_agents, level_text=level_text, subtasks=subtasks)
        cfg, self._keep = _cabi.make_config(
            self.level, num_envs=self.num_envs, num_agents=self.num_agents,
            max_num_timesteps=a.max_num_timesteps, num_communication=a.num_communication,
            communication_on=a.communication_on, ego_led=a.ego_led, fow_radius=a.fow_radius,
            ego_config=a.ego_config, partner_config=a.partner_config, seed=seed)
        self._handle = C.c_void_p()
        self._pinned = []
        self._closed = False
        self._pending = None
        self.lib.check(self.lib.set_device(self.device_index), "oc_set_device")
        self.lib.check(self.lib.create(C.byref(cfg), C.byref(self._handle)), "oc_create")
        self.obs_width = self.lib.obs_width(self._handle)
        off = (C.c_int32 * _cabi.OC_NUM_OBS_KEYS)()
        size = (C.c_int32 * _cabi.OC_NUM_OBS_KEYS)()
        self.lib.check(self.lib.obs_layout(self._handle, off, size), "oc_obs_layout")
        self.obs_layout = {k: slice(off[i], off[i] + size[i]) for i, k in enumerate(_cabi.OBS_KEYS)}
        self.observation_space, self.action_space = make_spaces(self.level, a.num_communication)
        E, A, F = self.num_envs, self.num_agents, self.obs_width
        i8 = obs_format == "i8"
        if i8 and self.obs_layout["timestep"] != slice(F - 1, F):
            raise RuntimeError("the compact format expects `timestep` to be the last key of a row")
        self._block_mode = False
        if i8:
            # one page-locked block for everything a step returns, when the handle has the compact-row kernels and
            # the messages fit a byte; otherwise the older entry points with separate buffers (same values)
            lay = _cabi.OcHostBlock()
            self.lib.check(self.lib.host_block_layout(self._handle, C.byref(lay)), "oc_host_block_layout")
            self._block_mode = a.num_communication <= 256 and self.lib.compact_supported(self._handle) == 1
        if self._block_mode:
            self._block = self.pinned_array((int(lay.total_bytes),), np.uint8)
            self.obs = self._block[lay.obs_i8:lay.obs_i8 + E * A * (F - 1)].view(np.int8).reshape(E, A, F - 1)
            self.timestep = self._block[lay.timestep:lay.timestep + 4 * E].view(np.float32)
            self.reward_per_env = self._block[lay.reward:lay.reward + 4 * E].view(np.float32)
            self.rewards = np.broadcast_to(self.reward_per_env[:, None], (E, A))         # read-only view
            self._dones_padded = self._block[lay.done:lay.done + (E + 7) // 8 * 8]         # sections are 256-byte padded
            self.actions_u8 = self.pinned_array((E, A, 2), np.uint8)
            self.actions = None
        else:
            odt, Fo = (np.int8, F - 1) if i8 else (np.float32, F)
            self.obs = self.pinned_array((E, A, Fo), odt)
            self.timestep = self.pinned_array((E,), np.float32) if i8 else None
            self.rewards = self.pinned_array((E, A), np.float32)
            self.reward_per_env = None
            self._dones_padded = self.pinned_array(((E + 7) // 8 * 8,), np.uint8)     # scanned as uint64 words
            self.actions = self.pinned_array((E, A, 2), np.int32)
            self.actions_u8 = None
        self.dones = self._dones_padded[:E]
        # what `step` takes without a conversion: [E, A, 2] of this dtype, ideally in `pinned_array` memory
        self.action_dtype = np.dtype(np.uint8 if self._block_mode else np.int32)
        odt, Fo = (np.int8, F - 1) if i8 else (np.float32, F)
        self.terminal_obs = self.pinned_array((E, A, Fo), odt) if (terminal_observations and auto_reset) else None
        self.terminal_timestep = self.pinned_array((E,), np.float32) if (i8 and self.terminal_obs is not None) else None
        self._infos = [{} for _ in range(E)]
        self._term_rows = [None] * E
        self._touched = ()
        # what crosses PCIe per step (bench.py's e2e keys)
        obs_bytes = self.obs.nbytes + (self.timestep.nbytes if self.timestep is not None else 0)
        if self._block_mode:
            self.h2d_bytes_per_step = self.actions_u8.nbytes
            self.d2h_bytes_per_step = int(lay.total_bytes)
            self.kernel_launches_per_step = 1
            self.transfer_desc = ("actions u8 [E,A,2] read by the step kernel from page-locked host memory; ONE device->host "
                                  "copy of the block obs_i8 | timestep | reward f32 [E] | done; terminal rows written by the kernel")
        else:
            self.h2d_bytes_per_step = self.actions.nbytes
            self.d2h_bytes_per_step = obs_bytes + self.rewards.nbytes + E
            self.kernel_launches_per_step = 1 if (not i8 or self.lib.compact_supported(self._handle) == 1) else 2
            self.transfer_desc = "actions int32 [E,A,2] host->device copy; separate device->host copies of obs, reward f32 [E,A], done"

    # ------------------------------------------------------------------ plumbing
    def _require_backend(self):
        if not isinstance(self.lib, _cabi.OcLibrary):
            raise RuntimeError("OvercookedHostVecEnv needs liboc_b200.so (CUDA); there is no CPU backend")

    def pinned_array(self, shape, dtype) -> np.ndarray:
        """numpy view of page-locked host memory from oc_host_alloc (freed in close()).  Arrays made here can be
        passed to `step` directly (no staging copy)."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = C.c_void_p()
        self.lib.check(self.lib.host_alloc(n, C.byref(ptr)), "oc_host_alloc")
        self._pinned.append(ptr)
        buf = (C.c_uint8 * max(n, 1)).from_address(ptr.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        arr[...] = 0
        return arr

    @staticmethod
    def _p(a: Optional[np.ndarray]):
        return None if a is None else C.c_void_p(a.ctypes.data)

    def _check(self, a: np.ndarray, shape, dtype, name):
        if not isinstance(a, np.ndarray) or a.dtype != dtype or tuple(a.shape) != tuple(shape) or not a.flags.c_contiguous:
            raise ValueError("%s must be a C-contiguous %s array of shape %s" % (name, np.dtype(dtype).name, tuple(shape)))

    # ------------------------------------------------------------------ API
    def reset(self, mask: Optional[np.ndarray] = None, placements: Optional[np.ndarray] = None) -> np.ndarray:
        if mask is not None:
            self._check(mask, (self.num_envs,), np.uint8, "mask")
        if placements is not None:
            self._check(placements, (self.num_envs, self.level.num_random), np.int32, "placements")
        self.lib.check(self.lib.set_device(self.device_index), "oc_set_device")
        if self._block_mode:
            self.lib.check(self.lib.reset_host_block(self._handle, self._p(mask), self._p(placements), self._p(self._block),
                                                     None), "oc_reset_host_block")
        elif self.obs_format == "i8":
            self.lib.check(self.lib.reset_host_i8(self._handle, self._p(mask), self._p(placements), self._p(self.obs),
                                                  self._p(self.timestep), None), "oc_reset_host_i8")
        else:
            self.lib.check(self.lib.reset_host(self._handle, self._p(mask), self._p(placements), self._p(self.obs), None),
                           "oc_reset_host")
        return self.obs

    def _stage_actions(self, actions):
        """-> the array handed to the C ABI: the caller's own buffer when it already has the wire format (uint8 pairs
        for the one-block path, int32 pairs otherwise), else the env's pinned buffer after one conversion."""
        want, buf = (np.uint8, self.actions_u8) if self._block_mode else (np.int32, self.actions)
        a = actions
        if isinstance(a, np.ndarray) and a.dtype == want and a.flags.c_contiguous and a.shape == buf.shape:
            return a
        a = np.asarray(actions)                      # any integer array-like: converted into the env's pinned buffer
        if a.shape != buf.shape:
            raise ValueError("actions must have shape %s (nav, comm per agent)" % (buf.shape,))
        np.copyto(buf, a, casting="unsafe" if self._block_mode else "same_kind")
        return buf

    def step_async(self, actions) -> None:
        """Enqueue the step (upload, kernel, download) and return; `step_wait` waits and builds the results.  On the
        one-block path the call does not block, so host work can overlap the GPU and the PCIe transfers."""
        a = self._stage_actions(actions)
        flags = _cabi.OC_FLAG_AUTO_RESET if self.auto_reset else 0
        self.lib.check(self.lib.set_device(self.device_index), "oc_set_device")
        if self._block_mode:
            self.lib.check(self.lib.step_host_block(self._handle, self._p(a), self._p(self._block), self._p(self.terminal_obs),
                                                    self._p(self.terminal_timestep), flags | _cabi.OC_FLAG_NO_SYNC, None),
                           "oc_step_host_block")
            self._pending = "block"
        else:
            self._pending = (a, flags)

    def step_wait(self):
        if self._pending is None:
            raise RuntimeError("step_wait without step_async")
        if self._pending == "block":
            self.lib.check(self.lib.sync(self._handle, None), "oc_sync")
        else:
            a, flags = self._pending
            if self.obs_format == "i8":
                self.lib.check(self.lib.step_host_i8(self._handle, self._p(a), self._p(self.obs), self._p(self.timestep),
                                                     self._p(self.rewards), None, self._p(self.dones),
                                                     self._p(self.terminal_obs), self._p(self.terminal_timestep), flags, None),
                               "oc_step_host_i8")
            else:
                self.lib.check(self.lib.step_host(self._handle, self._p(a), self._p(self.obs), self._p(self.rewards),
                                                  None, self._p(self.dones), self._p(self.terminal_obs), flags, None),
                               "oc_step_host")
        self._pending = None
        d = self.dones.view(np.bool_)
        # one dict per env; only the entries of envs that finished are touched
        infos = self._infos
        for e in self._touched:
            infos[e] = {}
        term = self.terminal_obs
        if term is None:
            self._touched = ()
            return self.obs, self.rewards, d, infos
        self._touched = idx = finished_indices(self._dones_padded, self.num_envs)
        rows = self._term_rows                       # views into the pinned buffer, made once per env
        if self.terminal_timestep is not None:
            for e, t in zip(idx, self.terminal_timestep[idx].tolist()):
                r = rows[e]
                if r is None:
                    r = rows[e] = term[e]
                infos[e] = {"terminal_observation": r, "terminal_timestep": t}
        else:
            for e in idx:
                r = rows[e]
                if r is None:
                    r = rows[e] = term[e]
                infos[e] = {"terminal_observation": r}
        return self.obs, self.rewards, d, self._infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    # packed state through host buffers (checkpointing; clock staggering for benchmarks)
    def get_state(self) -> np.ndarray:
        st = np.zeros((self.num_envs, _cabi.OC_STATE_WORDS), np.uint32)
        self.lib.check(self.lib.set_device(self.device_index), "oc_set_device")
        self.lib.check(self.lib.get_state_host(self._handle, self._p(st), None), "oc_get_state_host")
        return st

    def set_state(self, st: np.ndarray) -> None:
        self._check(st, (self.num_envs, _cabi.OC_STATE_WORDS), np.uint32, "state")
        self.lib.check(self.lib.set_device(self.device_index), "oc_set_device")
        self.lib.check(self.lib.set_state_host(self._handle, self._p(st), None), "oc_set_state_host")

    def stagger_clocks(self, period: Optional[int] = None, multiplier: int = 1) -> None:
        """Env e's episode clock := (e * multiplier) mod period (default: max_num_timesteps, 1): the steady state of a
        long run, in which about E / T envs finish in every step instead of all of them in the same one; a multiplier
        coprime to the period spreads the envs that finish in one step over the batch."""
        T = int(period or self.arglist.max_num_timesteps)
        st = self.get_state()
        clocks = (np.arange(self.num_envs, dtype=np.uint64) * np.uint64(multiplier)) % np.uint64(T)
        st[:, 0] = (st[:, 0] & np.uint32(0xFFFF0000)) | clocks.astype(np.uint32)
        self.set_state(st)

    # SB3 VecEnv duck type (the rest of the convention; every env shares one configuration)
    def seed(self, seed=None):
        return [None] * self.num_envs          # placements are seeded at construction (oc_config.seed)

    def get_attr(self, attr_name, indices=None):
        n = self.num_envs if indices is None else len([indices] if isinstance(indices, int) else list(indices))
        return [getattr(self, attr_name)] * n

    def env_is_wrapped(self, wrapper_class, indices=None):
        n = self.num_envs if indices is None else len([indices] if isinstance(indices, int) else list(indices))
        return [False] * n

    def obs_dict(self, obs: Optional[np.ndarray] = None, timestep: Optional[np.ndarray] = None) -> dict:
        """Per-key zero-copy views of flat rows (the reference's Dict observation).  Compact format: the
        integer keys are int8 views and `timestep` is the per-env clock broadcast to [E, A, 1]."""
        o = self.obs if obs is None else obs
        if self.obs_format != "i8":
            return {k: o[..., s] for k, s in self.obs_layout.items()}
        t = self.timestep if timestep is None else timestep
        d = {k: o[..., s] for k, s in self.obs_layout.items() if k != "timestep"}
        d["timestep"] = np.broadcast_to(np.asarray(t, np.float32).reshape(-1, 1, 1), o.shape[:-1] + (1,))
        return d

    def obs_float(self, obs: Optional[np.ndarray] = None, timestep: Optional[np.ndarray] = None) -> np.ndarray:
        """The float32 rows [E, A, F] (a new array) -- what `obs_format="f32"` returns directly."""
        o = self.obs if obs is None else obs
        if self.obs_format != "i8":
            return np.array(o, np.float32)
        t = self.timestep if timestep is None else timestep
        out = np.empty(o.shape[:-1] + (self.obs_width,), np.float32)
        out[..., :-1] = o
        out[..., -1] = np.asarray(t, np.float32).reshape(-1, 1)
        return out

    def close(self):
        if self._closed:
            return
        self._closed = True
        try:
            self.lib.set_device(self.device_index)
            if self._handle:
                self.lib.destroy(self._handle)
            for ptr in self._pinned:
                self.lib.host_free(ptr)
        finally:
            self._handle = C.c_void_p()
            self._pinned = []
            self.obs = self.rewards = self.dones = self._dones_padded = self.actions = self.terminal_obs = None
            self.actions_u8 = self.reward_per_env = self._block = None
            self._term_rows = []
            self.timestep = self.terminal_timestep = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
