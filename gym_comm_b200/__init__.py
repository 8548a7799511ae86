"""gym_comm_b200 -- B200-native batched Overcooked simulator (drop-in for the env step +
observation path of kyle-he/gym-comm).  See DESIGN.md."""
from .arglist import create_arglist, namespace_from_dict  # noqa: F401
from .level_compiler import compile_level  # noqa: F401
from .vec_env import OvercookedMultiEnv, OvercookedVecEnv  # noqa: F401

__all__ = ["OvercookedVecEnv", "OvercookedMultiEnv", "create_arglist", "namespace_from_dict", "compile_level"]
