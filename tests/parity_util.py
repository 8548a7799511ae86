"""Shared parity machinery: replay a golden trace (recorded from the live reference) through an
``OvercookedVecEnv`` and compare EVERYTHING bit-exactly each step: returned reward (f64), done,
every observer's flat observation, and the packed state decoded back to the reference's terms.

Used with the real CUDA library (tests marked gpu) and with tests/emu (the same device source
executed on the CPU, for debugging in the GPU-less build container)."""
import argparse
import os
import subprocess

import numpy as np
import torch

from gym_comm_b200 import _cabi
from gym_comm_b200.vec_env import OvercookedMultiEnv, OvercookedVecEnv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")


EMU_EXPORTS = ("oc_last_error", "oc_create", "oc_destroy", "oc_obs_width", "oc_obs_layout", "oc_reset", "oc_step",
               "oc_reset_i8", "oc_step_i8", "oc_rollout", "oc_get_state", "oc_set_state", "oc_get_stats", "oc_pack_obs_i8")


class EmuLibrary:
    """TEST ONLY.  The CPU emulation of the device code (tests/emu/liboc_emu.so, symbols `emu_*`) behind the
    attribute surface of `_cabi.OcLibrary`, so the host-side Python of the product can be exercised in the
    GPU-less build container.  The product package knows nothing about it: `EmuVecEnv` below is the only way in."""

    def __init__(self, path):
        import ctypes as C
        self.path = path
        self.lib = C.CDLL(path)
        for name in EMU_EXPORTS:
            f = getattr(self.lib, "emu_" + name[3:])
            f.restype, f.argtypes = _cabi.SIGNATURES[name]
            setattr(self, name[3:], f)

    def check(self, rc, what):
        if rc != 0:
            msg = self.last_error()
            raise RuntimeError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))

    def launch_count(self, handle):
        return 0


_emu = None


def emu_library():
    """Build (if stale) and load the CPU emulation of the device code.  TEST ONLY."""
    global _emu
    so = os.path.join(EMU_DIR, "liboc_emu.so")
    srcs = [os.path.join(EMU_DIR, "oc_emu.cpp"), os.path.join(EMU_DIR, "oc_emu_shim.h")] + \
           [os.path.join(ROOT, "gym_comm_b200", "csrc", f) for f in ("oc_device.cuh", "oc_host.hpp", "oc_params.h")] + \
           [os.path.join(ROOT, "include", "overcooked_b200.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I" + EMU_DIR,
                               os.path.join(EMU_DIR, "oc_emu.cpp"), "-o", so])
        _emu = None
    if _emu is None:
        _emu = EmuLibrary(so)
    return _emu


class EmuVecEnv(OvercookedVecEnv):
    """TEST ONLY: `OvercookedVecEnv` with its CUDA requirement lifted and the emulation library behind it
    (CPU tensors, no stream).  Lives here, not in the product, so that the product has no CPU path."""

    def __init__(self, arglist, num_envs=1, device="cpu", lib=None, **kw):
        super().__init__(arglist, num_envs=num_envs, device="cpu", lib=lib if lib is not None else emu_library(), **kw)

    def _require_backend(self):
        assert isinstance(self.lib, EmuLibrary)

    def _device_guard(self):
        import contextlib
        return contextlib.nullcontext()

    def _stream(self):
        return None


class EmuMultiEnv(OvercookedMultiEnv):
    """TEST ONLY: the reference-shaped 2-player surface over `EmuVecEnv`."""
    _vec_cls = EmuVecEnv


def namespace_from_meta(meta):
    return argparse.Namespace(
        level=meta["level"], num_agents=meta["num_agents"], max_num_timesteps=meta["max_num_timesteps"],
        communication_on=meta["communication_on"], num_communication=meta["num_communication"],
        ego_led=meta["ego_led"], fow_radius=meta["fow_radius"], ego_config=meta["ego_config"],
        partner_config=meta["partner_config"])


def world_order_objects(dec, e, level):
    """Alive objects of env e in world.objects iteration order -> rows like the golden dump."""
    rows = []
    ranks = int(dec["ranks"][e])
    for s in range(6):
        c = int(dec["obj_contents"][e, s])
        if c == 0:
            continue
        rank = (ranks >> (4 * c)) & 15
        cell = int(dec["obj_cell"][e, s])
        rows.append(((rank, int(dec["obj_stamp"][e, s])),
                     [c, int(dec["obj_chopped"][e, s]), cell % level.width, cell // level.width,
                      int(dec["obj_holder"][e, s] != 7)]))
    rows.sort(key=lambda r: r[0])
    return [r[1] for r in rows]


def replay_golden(meta, g, lib, device, num_envs=3):
    ns = namespace_from_meta(meta)
    cls = EmuVecEnv if isinstance(lib, EmuLibrary) else OvercookedVecEnv
    env = cls(ns, num_envs=num_envs, device=device, auto_reset=False, lib=lib,
              level_text=meta["level_text"], subtasks=meta["subtasks"])
    lv = env.level
    n, E = meta["num_agents"], num_envs
    S = len(meta["subtasks"])

    def placements(ep):
        pl = g["placements"][ep]
        if pl.shape[0] == 0:
            return None
        cells = np.array([lv.cell(int(x), int(y)) for x, y in pl], dtype=np.int32)
        return torch.from_numpy(np.tile(cells, (E, 1))).to(device)

    def check_obs(obs, want, where):
        got = obs.cpu().numpy()
        want32 = want.astype(np.float32)
        for e in range(E):
            assert np.array_equal(got[e], want32), (meta["name"], where, e, np.argwhere(got[e] != want32)[:5],
                                                    got[e][got[e] != want32][:5], want32[got[e] != want32][:5])

    ep = 0
    check_obs(env.reset(placements=placements(0)), g["reset_obs"][0], "reset0")
    acts = torch.zeros((E, n, 2), dtype=torch.int32, device=device)
    for i in range(len(g["navs"])):
        a = np.stack([g["navs"][i].astype(np.int32), g["comms"][i].astype(np.int32)], -1)   # [n, 2]
        acts.copy_(torch.from_numpy(np.tile(a, (E, 1, 1))))
        obs, rew, done = env.step(acts, want_f64=True)
        r64 = env.rewards64.cpu().numpy()
        assert np.all(r64 == g["reward"][i]), (meta["name"], i, r64, g["reward"][i])
        assert np.all(rew.cpu().numpy() == np.float32(g["reward"][i])), (meta["name"], i)
        assert np.all(done.cpu().numpy() == int(g["done"][i])), (meta["name"], i)
        check_obs(obs, g["obs"][i], i)
        dec = env.decode_state()
        for e in range(E):
            assert int(dec["t"][e]) == int(g["t"][i])
            assert [int((dec["completed"][e] >> j) & 1) for j in range(S)] == g["completed"][i].tolist(), (meta["name"], i)
            deliver = [k == 2 for k in lv.subtask_kind]
            want_counts = [0 if deliver[j] else int(g["counts"][i][j]) for j in range(S)]
            assert [int((dec["countbits"][e] >> j) & 1) for j in range(S)] == want_counts, (meta["name"], i)
            assert np.array_equal(np.stack([dec["agent_x"][e], dec["agent_y"][e]], -1), g["agents"][i]), (meta["name"], i)
            want_objs = [row for row in g["objs"][i].tolist() if row[0] >= 0]
            assert world_order_objects(dec, e, lv) == want_objs, (meta["name"], i, world_order_objects(dec, e, lv), want_objs)
        if g["done"][i]:
            ep += 1
            check_obs(env.reset(placements=placements(ep)), g["reset_obs"][ep], ("reset", ep))
    env.close()


class EmuHostLibrary(_cabi.OcLibrary):
    """TEST ONLY.  Stands where `liboc_b200.so` stands for `OvercookedHostVecEnv`, so that the Python side of
    the host-buffer path (buffers, formats, infos, terminal observations, the one-block layout) can run in the
    GPU-less container: the host-buffer entry points are re-expressed over the CPU emulation of the device
    functions (emu_step / emu_step_i8 / emu_reset_i8 -- the kernels the real entry points launch).  The real
    entry points are covered on the GPU by tests/test_gpu_host_env.py and tests/cabi_smoke.c.
    (Subclasses OcLibrary only to pass the product's isinstance check; nothing of the CUDA library is loaded.)"""

    def __init__(self, compact=True):
        import ctypes as C
        self.C = C
        self.emu = emu_library()
        self.destroy = self.emu.destroy
        self.obs_width, self.obs_layout = self.emu.obs_width, self.emu.obs_layout
        self.last_error = self.emu.last_error
        self._gather = self.emu.lib.emu_gather_term
        self._gather.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3
        self._bufs, self._dims = {}, {}
        self._compact = compact

    def check(self, rc, what):
        assert rc == 0, (what, rc)

    def set_device(self, index):
        return 0

    def sync(self, h, stream):
        return 0

    def compact_supported(self, h):
        return 1 if self._compact else 0

    def host_alloc(self, n, ref):
        buf = (self.C.c_uint8 * max(int(n), 1))()
        addr = self.C.addressof(buf)
        self._bufs[addr] = buf
        ref._obj.value = addr
        return 0

    def host_free(self, ptr):
        self._bufs.pop(ptr.value, None)
        return 0

    def create(self, cfg, href):
        rc = self.emu.create(cfg, href)
        if rc == 0:                      # remember (E, A, F) of the handle (the C library knows them from oc_config)
            self._dims[href._obj.value] = (cfg._obj.num_envs, cfg._obj.num_agents, self.emu.obs_width(href._obj))
        return rc

    def bind(self, env):
        pass

    @staticmethod
    def _a(ptr, shape, dtype):
        import ctypes as C
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        return np.frombuffer((C.c_uint8 * n).from_address(ptr.value), dtype=dtype).reshape(shape)

    def get_state_host(self, h, st, stream):
        return self.emu.get_state(h, st, None)

    def set_state_host(self, h, st, stream):
        return self.emu.set_state(h, st, None)

    def reset_host(self, h, mask, placements, obs, stream):
        return self.emu.reset(h, mask, placements, obs, None)

    def step_host(self, h, actions, obs, rew, rew64, done, term, flags, stream):
        return self.emu.step(h, actions, obs, rew, rew64, done, term, flags, None)

    def reset_host_i8(self, h, mask, placements, obs8, ts, stream):
        if self._compact:
            return self.emu.reset_i8(h, mask, placements, obs8, ts, None)
        E, A, F = self._dims[h.value]
        f = np.zeros((E, A, F), np.float32)
        rc = self.emu.reset(h, mask, placements, self.C.c_void_p(f.ctypes.data), None)
        return rc or self.emu.pack_obs_i8(h, self.C.c_void_p(f.ctypes.data), obs8, ts, None)

    def step_host_i8(self, h, actions, obs8, ts, rew, rew64, done, term8, term_ts, flags, stream):
        C = self.C
        if self._compact:                 # the compact-row kernel writes everything itself, terminal rows included
            return self.emu.step_i8(h, actions, obs8, ts, rew, rew64, done, term8, term_ts, flags, None)
        E, A, F = self._dims[h.value]     # rows too wide for the compact kernels: float rows, repack, gather
        f = np.zeros((E, A, F), np.float32)
        t = np.zeros((E, A, F), np.float32)
        rc = self.emu.step(h, actions, C.c_void_p(f.ctypes.data), rew, rew64, done,
                           C.c_void_p(t.ctypes.data) if term8 is not None else None, flags, None)
        rc = rc or self.emu.pack_obs_i8(h, C.c_void_p(f.ctypes.data), obs8, ts, None)
        if rc or term8 is None:
            return rc
        idx = np.flatnonzero(self._a(done, (E,), np.uint8)).astype(np.int32)
        if idx.size:
            g8, gts = np.zeros((idx.size, A, F - 1), np.int8), np.zeros(idx.size, np.float32)
            self._gather(h, t.ctypes.data, idx.ctypes.data, idx.size, None, g8.ctypes.data, gts.ctypes.data)
            self._a(term8, (E, A, F - 1), np.int8)[idx] = g8
            if term_ts is not None:
                self._a(term_ts, (E,), np.float32)[idx] = gts
        return 0

    # the one-block path: same layout rule as oc_host_block_layout (sections padded to 256 bytes)
    def host_block_layout(self, h, ref):
        E, A, F = self._dims[h.value]
        lay = ref._obj
        off = 0
        for name, n in (("obs_i8", E * A * (F - 1)), ("timestep", 4 * E), ("reward", 4 * E), ("done", E)):
            setattr(lay, name, off)
            off += (n + 255) // 256 * 256
        lay.total_bytes = off
        return 0

    def _views(self, h, block):
        E, A, F = self._dims[h.value]
        lay = _cabi.OcHostBlock()
        self.host_block_layout(h, self.C.byref(lay))
        base = block.value
        p = lambda off: self.C.c_void_p(base + off)
        return p(lay.obs_i8), p(lay.timestep), p(lay.reward), p(lay.done)

    def reset_host_block(self, h, mask, placements, block, stream):
        obs8, ts, _, _ = self._views(h, block)
        return self.emu.reset_i8(h, mask, placements, obs8, ts, None)

    def step_host_block(self, h, actions_u8, block, term8, term_ts, flags, stream):
        obs8, ts, rew, done = self._views(h, block)
        fl = (flags & _cabi.OC_FLAG_AUTO_RESET) | _cabi.OC_FLAG_ACTIONS_U8 | _cabi.OC_FLAG_REWARD_PER_ENV
        return self.emu.step_i8(h, actions_u8, obs8, ts, rew, None, done, term8, term_ts, fl, None)
