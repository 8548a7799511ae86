"""CPU-side debug aid: the device source (oc_device.cuh) compiled for the host and executed lane by
lane (tests/emu) must reproduce the golden traces.  This does NOT count as GPU parity -- the
tests marked ``gpu`` run the real kernels through the C ABI -- but it catches logic bugs in the
packed-state code in a container without a GPU."""
import pytest

from tests.golden_util import golden_names, load_golden
from tests.parity_util import emu_library, replay_golden


@pytest.fixture(scope="module")
def emu():
    return emu_library()


@pytest.mark.parametrize("name", golden_names())
def test_device_code_on_cpu_replays_golden(emu, name):
    meta, g = load_golden(name)
    replay_golden(meta, g, emu, "cpu", num_envs=3 if meta["num_communication"] < 50 else 33)
