"""The C-ABI library loads on a CPU-only box and exports every symbol include/overcooked_b200.h
declares; without a CUDA device creation fails loudly (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build_cuda()
    from gym_comm_b200 import _cabi
    return _cabi.default_library()


def test_exports_match_header(lib):
    from gym_comm_b200 import _cabi
    hdr = open(os.path.join(ROOT, "include", "overcooked_b200.h")).read()
    declared = set(re.findall(r"^(?:int|uint64_t|const char\*)\s+(oc_\w+)\s*\(", hdr, re.M))
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)
    for name in declared:
        assert hasattr(lib.lib, name), name
    assert lib.abi_version() == _cabi.OC_ABI_VERSION


def test_config_struct_size_matches_header():
    """ctypes mirror of oc_config must have the C layout (checked with a tiny gcc probe)."""
    import subprocess
    import tempfile
    from gym_comm_b200 import _cabi
    import ctypes
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "overcooked_b200.h"\nint main(){printf("%zu %zu %zu %zu", sizeof(oc_config), offsetof(oc_config, tiles), offsetof(oc_config, subtask_kind), offsetof(oc_config, seed));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "p.c"), "-o", os.path.join(d, "p")])
        out = subprocess.check_output([os.path.join(d, "p")]).decode().split()
    c = _cabi.OcConfig
    assert [int(v) for v in out] == [ctypes.sizeof(c), c.tiles.offset, c.subtask_kind.offset, c.seed.offset]


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import argparse
    from gym_comm_b200.vec_env import OvercookedVecEnv
    ns = argparse.Namespace(level="open-divider_tomato", num_agents=2, max_num_timesteps=100)
    with pytest.raises(RuntimeError):
        OvercookedVecEnv(ns, num_envs=4, device="cuda")
    with pytest.raises(RuntimeError):
        OvercookedVecEnv(ns, num_envs=4, device="cpu")


def test_host_env_has_no_cpu_fallback_either(lib):
    """The numpy-only host path needs the CUDA library and a device; it refuses the test emulation."""
    import torch
    import argparse
    from gym_comm_b200.host_env import OvercookedHostVecEnv
    from tests.parity_util import emu_library
    ns = argparse.Namespace(level="open-divider_tomato", num_agents=2, max_num_timesteps=100)
    with pytest.raises(RuntimeError):
        OvercookedHostVecEnv(ns, num_envs=4, lib=emu_library())
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            OvercookedHostVecEnv(ns, num_envs=4)


def test_emulation_backend_is_refused_outside_tests(monkeypatch):
    import argparse
    from gym_comm_b200.vec_env import OvercookedVecEnv
    from tests.parity_util import emu_library
    emu = emu_library()
    monkeypatch.delenv("OC_TEST_EMULATION", raising=False)
    ns = argparse.Namespace(level="open-divider_tomato", num_agents=2, max_num_timesteps=100)
    with pytest.raises(RuntimeError):
        OvercookedVecEnv(ns, num_envs=4, device="cpu", lib=emu)
