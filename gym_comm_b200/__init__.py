"""gym_comm_b200 -- B200-native batched Overcooked simulator (drop-in for the env step +
observation path of kyle-he/gym-comm).  See DESIGN.md."""
from .arglist import create_arglist, namespace_from_dict  # noqa: F401
from .level_compiler import compile_level  # noqa: F401

__all__ = ["OvercookedVecEnv", "OvercookedMultiEnv", "OvercookedHostVecEnv", "create_arglist", "namespace_from_dict",
           "compile_level"]


def __getattr__(name):
    # the torch-based classes are imported on first use so that the numpy-only host path
    # (OvercookedHostVecEnv) does not pull torch in
    if name in ("OvercookedVecEnv", "OvercookedMultiEnv"):
        from . import vec_env
        return getattr(vec_env, name)
    if name == "OvercookedHostVecEnv":
        from .host_env import OvercookedHostVecEnv
        return OvercookedHostVecEnv
    raise AttributeError(name)
