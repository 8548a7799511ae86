"""The compact-row kernels' device code (MODE 3 of the step / reset kernels) on the CPU emulation: see
tests/compact_cases.py.  The CUDA kernels themselves run the same cases in tests/test_gpu_compact.py."""
import pytest

from tests import compact_cases as cases
from tests.parity_util import EmuVecEnv


def _make(ns, **kw):
    return EmuVecEnv(ns, device="cpu", **kw)


@pytest.mark.parametrize("level,A,T,C,fow", cases.CASES)
@pytest.mark.parametrize("E,u8,per_env", [(75, False, False), (33, True, True)])
def test_step_i8_equals_float_rows(level, A, T, C, fow, E, u8, per_env):
    cases.run_step_i8_equals_float_rows(_make, "cpu", level, A, T, C, fow, E, u8, per_env)


def test_set_state_sanitises():
    cases.run_set_state_sanitises(_make, "cpu")
