"""GPU parity, part 1: the real sm_100a kernels, called through the C ABI, replay every golden
trace recorded from the live reference -- reward (exact f64 and its f32 rounding), done, all
observations, and the packed state -- bit-exactly, for a batch of identical envs."""
import pytest

from tests.golden_util import golden_names, load_golden
from tests.parity_util import replay_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from gym_comm_b200 import _cabi
    return _cabi.default_library()


@pytest.mark.parametrize("name", golden_names())
def test_cuda_replays_golden(lib, name):
    meta, g = load_golden(name)
    replay_golden(meta, g, lib, "cuda:0", num_envs=70)
