"""gym-free stand-ins for the observation / action spaces the reference declares
(gym_comm/envs/overcooked_env.py:57-85); gym itself is not required."""
from __future__ import annotations

import numpy as np

from .level_compiler import NAV_ACTIONS


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high, self.dtype = low, high, dtype
        self.shape = tuple(shape) if shape is not None else np.asarray(low).shape


class MultiBinary:
    def __init__(self, n):
        self.n, self.shape, self.dtype = n, (n,), np.int8


class MultiDiscrete:
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec)
        self.shape, self.dtype = self.nvec.shape, np.int64


class Dict:
    def __init__(self, spaces):
        # gym.spaces.Dict sorts a plain dict by key; that order is the flat feature order
        self.spaces = dict(sorted(spaces.items()))


def make_spaces(level, num_communication: int):
    """-> (observation_space, action_space) of one agent (overcooked_env.py:57-85)."""
    S, Cn = len(level.subtasks), int(num_communication)
    w, h = level.width, level.height
    obs = Dict({
        "timestep": Box(0.0, 1.0, (1,), np.float32),
        "object_encodings_x": Box(-w, w, (4,), np.int64),
        "object_encodings_y": Box(h, h, (4,), np.int64),          # sic: low == high in the reference (:62)
        "state_encodings": MultiBinary(4), "is_hidden": MultiBinary(4),
        "completed_subtasks": MultiBinary(S),
        "agent1_location": Box(np.array([0, 0]), np.array([w - 1, h - 1]), dtype=np.float32),
        "agent2_location": Box(np.array([0, 0]), np.array([w - 1, h - 1]), dtype=np.float32),
        "agent_is_holding": MultiBinary(2),
        "agent1_comm": MultiBinary(Cn), "agent2_comm": MultiBinary(Cn)})
    return obs, MultiDiscrete([len(NAV_ACTIONS), Cn])      # overcooked_env.py:85
