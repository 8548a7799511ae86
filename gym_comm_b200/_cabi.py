"""ctypes binding of the C ABI declared in include/overcooked_b200.h.

The product path is ``liboc_b200.so`` (hand-written sm_100a CUDA, built in-tree by
``__graft_entry__.build()``).  There is no CPU fallback: if the library is missing this module
raises, and ``oc_create`` fails when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

OC_ABI_VERSION = 2
OC_MAX_AGENTS, OC_MAX_OBJECTS, OC_MAX_SUBTASKS, OC_MAX_CELLS = 4, 6, 32, 128
OC_STATE_WORDS = 16
OC_NUM_OBS_KEYS = 11
OC_FLAG_AUTO_RESET = 1
OC_FLAG_ACTIONS_U8 = 2
OC_FLAG_REWARD_PER_ENV = 4
OC_FLAG_NO_SYNC = 8
OC_FLAG_CHAIN_HEAD = 16
OC_FLAG_CHAINED = 32
OBS_KEYS = ("agent1_comm", "agent1_location", "agent2_comm", "agent2_location", "agent_is_holding",
            "completed_subtasks", "is_hidden", "object_encodings_x", "object_encodings_y",
            "state_encodings", "timestep")

EXPORTS = ("oc_abi_version", "oc_last_error", "oc_create", "oc_destroy", "oc_obs_width", "oc_obs_layout",
           "oc_reset", "oc_step", "oc_rollout", "oc_replay", "oc_get_state", "oc_set_state", "oc_get_stats",
           "oc_launch_count", "oc_reset_host", "oc_step_host", "oc_pack_obs_i8", "oc_reset_host_i8",
           "oc_step_host_i8", "oc_host_alloc", "oc_host_free", "oc_set_device", "oc_reset_i8", "oc_step_i8",
           "oc_host_block_layout", "oc_reset_host_block", "oc_step_host_block", "oc_sync", "oc_get_state_host",
           "oc_set_state_host", "oc_compact_supported")


class OcConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("num_envs", C.c_int32),
        ("num_agents", C.c_int32),
        ("width", C.c_int32), ("height", C.c_int32),
        ("max_num_timesteps", C.c_int32),
        ("num_communication", C.c_int32),
        ("communication_on", C.c_int32),
        ("ego_led", C.c_int32),
        ("fow_radius", C.c_int32),
        ("can_move", C.c_uint8 * OC_MAX_AGENTS),
        ("allergic", C.c_uint8 * OC_MAX_AGENTS),
        ("blind", C.c_uint8 * OC_MAX_AGENTS),
        ("start_cell", C.c_uint8 * OC_MAX_AGENTS),
        ("tiles", C.POINTER(C.c_uint8)),
        ("path_dist", C.POINTER(C.c_uint8)),
        ("max_path", C.c_int32),
        ("num_objects", C.c_int32),
        ("object_contents", C.c_uint8 * OC_MAX_OBJECTS),
        ("object_cell", C.c_int16 * OC_MAX_OBJECTS),
        ("num_subtasks", C.c_int32),
        ("subtask_kind", C.c_uint8 * OC_MAX_SUBTASKS),
        ("subtask_goal", C.c_uint8 * OC_MAX_SUBTASKS),
        ("subtask_arg0", C.c_uint8 * OC_MAX_SUBTASKS),
        ("num_items", C.c_int32),
        ("items", C.c_uint8 * 4),
        ("seed", C.c_uint64),
    ]


DEFAULT_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liboc_b200.so")


class OcHostBlock(C.Structure):
    """oc_host_block: byte offsets of the sections of the one-block host path."""
    _fields_ = [("obs_i8", C.c_uint64), ("timestep", C.c_uint64), ("reward", C.c_uint64), ("done", C.c_uint64),
                ("total_bytes", C.c_uint64)]


def _signatures():
    vp, i32, u32 = C.c_void_p, C.c_int32, C.c_uint32
    return {
        "oc_last_error": (C.c_char_p, []),
        "oc_create": (C.c_int, [C.POINTER(OcConfig), C.POINTER(vp)]),
        "oc_destroy": (C.c_int, [vp]),
        "oc_obs_width": (C.c_int, [vp]),
        "oc_obs_layout": (C.c_int, [vp, C.POINTER(i32), C.POINTER(i32)]),
        "oc_reset": (C.c_int, [vp, vp, vp, vp, vp]),
        "oc_step": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, u32, vp]),
        "oc_reset_i8": (C.c_int, [vp, vp, vp, vp, vp, vp]),
        "oc_step_i8": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, u32, vp]),
        "oc_rollout": (C.c_int, [vp, i32, vp, vp, vp, vp, vp]),
        "oc_replay": (C.c_int, [vp, i32, vp, vp, vp, vp, vp]),
        "oc_get_state": (C.c_int, [vp, vp, vp]),
        "oc_set_state": (C.c_int, [vp, vp, vp]),
        "oc_get_stats": (C.c_int, [vp, vp, vp, vp]),
        "oc_pack_obs_i8": (C.c_int, [vp, vp, vp, vp, vp]),
        "oc_reset_host": (C.c_int, [vp, vp, vp, vp, vp]),
        "oc_step_host": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, u32, vp]),
        "oc_reset_host_i8": (C.c_int, [vp, vp, vp, vp, vp, vp]),
        "oc_step_host_i8": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, u32, vp]),
        "oc_host_block_layout": (C.c_int, [vp, C.POINTER(OcHostBlock)]),
        "oc_reset_host_block": (C.c_int, [vp, vp, vp, vp, vp]),
        "oc_step_host_block": (C.c_int, [vp, vp, vp, vp, vp, u32, vp]),
        "oc_sync": (C.c_int, [vp, vp]),
        "oc_get_state_host": (C.c_int, [vp, vp, vp]),
        "oc_set_state_host": (C.c_int, [vp, vp, vp]),
        "oc_host_alloc": (C.c_int, [C.c_uint64, C.POINTER(vp)]),
        "oc_host_free": (C.c_int, [vp]),
        "oc_set_device": (C.c_int, [C.c_int]),
        "oc_abi_version": (C.c_int, []),
        "oc_launch_count": (C.c_uint64, [vp]),
        "oc_compact_supported": (C.c_int, [vp]),
    }


SIGNATURES = _signatures()      # (restype, argtypes) of every export of include/overcooked_b200.h


class OcLibrary:
    """`liboc_b200.so` (hand-written sm_100a CUDA) + typed entry points.  This is the only backend the package
    has: there is no CPU implementation to select."""

    def __init__(self, path: str = DEFAULT_LIB):
        if not os.path.exists(path):
            raise RuntimeError(
                "%s not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). There is no CPU fallback." % path)
        self.path = path
        self.lib = C.CDLL(path)
        for name in EXPORTS:                       # self.step = oc_step, self.reset_host_i8 = oc_reset_host_i8, ...
            f = getattr(self.lib, name)
            f.restype, f.argtypes = SIGNATURES[name]
            setattr(self, name[3:], f)
        if self.abi_version() != OC_ABI_VERSION:
            raise RuntimeError("liboc_b200.so ABI version %d != %d (rebuild: __graft_entry__.build())" %
                               (self.abi_version(), OC_ABI_VERSION))

    def check(self, rc: int, what: str):
        if rc != 0:
            msg = self.last_error()
            raise RuntimeError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))


_default = None


def default_library() -> OcLibrary:
    """The in-tree CUDA library.  `OC_B200_LIB=<path>` loads another BUILD of the same sources instead (A/B of
    compile-time knobs, e.g. the -DOC_PHASE_PROBE build of tools/probe_step.py); it is still the CUDA library."""
    global _default
    if _default is None:
        _default = OcLibrary(os.environ.get("OC_B200_LIB") or DEFAULT_LIB)
    return _default


def make_config(level, *, num_envs, num_agents, max_num_timesteps, num_communication, communication_on,
                ego_led, fow_radius, ego_config, partner_config, seed=0):
    """CompiledLevel + arglist fields -> (OcConfig, keepalive) .  Agent 0 gets ``ego_config``, every
    other agent ``partner_config`` (overcooked_environment.py:140-143)."""
    import numpy as np
    cfg = OcConfig()
    cfg.abi_version = OC_ABI_VERSION
    cfg.num_envs = int(num_envs)
    cfg.num_agents = int(num_agents)
    cfg.width, cfg.height = level.width, level.height
    cfg.max_num_timesteps = int(max_num_timesteps)
    cfg.num_communication = int(num_communication)
    cfg.communication_on = int(bool(communication_on))
    cfg.ego_led = int(bool(ego_led))
    cfg.fow_radius = int(fow_radius)
    for k in range(num_agents):
        ac = ego_config if k == 0 else partner_config
        cfg.can_move[k] = int(bool(ac["CAN_MOVE"]))
        cfg.allergic[k] = int(bool(ac["ALLERGIC"]))
        cfg.blind[k] = int(bool(ac["BLIND"]))
        cfg.start_cell[k] = level.starts[k]
    tiles = np.ascontiguousarray(level.tiles, dtype=np.uint8)
    pd = np.ascontiguousarray(level.path_dist, dtype=np.uint8)
    cfg.tiles = tiles.ctypes.data_as(C.POINTER(C.c_uint8))
    cfg.path_dist = pd.ctypes.data_as(C.POINTER(C.c_uint8))
    cfg.max_path = level.max_path
    cfg.num_objects = len(level.object_contents)
    for s, (b, c) in enumerate(zip(level.object_contents, level.object_cell)):
        cfg.object_contents[s] = b
        cfg.object_cell[s] = c
    cfg.num_subtasks = len(level.subtasks)
    for i in range(len(level.subtasks)):
        cfg.subtask_kind[i] = level.subtask_kind[i]
        cfg.subtask_goal[i] = level.subtask_goal[i]
        cfg.subtask_arg0[i] = level.subtask_arg0[i] if level.subtask_kind[i] == 0 else 0
    cfg.num_items = len(level.items)
    for i, b in enumerate(level.items):
        cfg.items[i] = b
    cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return cfg, (tiles, pd)
