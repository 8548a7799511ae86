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
from gym_comm_b200.vec_env import OvercookedVecEnv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")


def emu_library():
    """Build (if stale) and load the CPU emulation of the device code.  TEST ONLY."""
    so = os.path.join(EMU_DIR, "liboc_emu.so")
    srcs = [os.path.join(EMU_DIR, "oc_emu.cpp"), os.path.join(EMU_DIR, "oc_emu_shim.h")] + \
           [os.path.join(ROOT, "gym_comm_b200", "csrc", f) for f in ("oc_device.cuh", "oc_host.hpp", "oc_params.h")] + \
           [os.path.join(ROOT, "include", "overcooked_b200.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I" + EMU_DIR,
                               os.path.join(EMU_DIR, "oc_emu.cpp"), "-o", so])
    os.environ["OC_TEST_EMULATION"] = "1"          # the product refuses non-CUDA backends without this
    return _cabi.OcLibrary(so, prefix="emu_")


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
    env = OvercookedVecEnv(ns, num_envs=num_envs, device=device, auto_reset=False, lib=lib,
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
