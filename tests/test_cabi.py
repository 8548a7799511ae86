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


def test_host_env_refuses_anything_but_the_cuda_library(lib):
    """The numpy host path has no CPU fallback either: only an `OcLibrary` (liboc_b200.so) is accepted, and
    without a device creation fails."""
    import torch
    import argparse
    from gym_comm_b200.host_env import OvercookedHostVecEnv
    from tests.parity_util import emu_library
    ns = argparse.Namespace(level="open-divider_tomato", num_agents=2, max_num_timesteps=100)
    with pytest.raises(RuntimeError, match="no CPU backend"):
        OvercookedHostVecEnv(ns, num_envs=4, lib=emu_library())
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            OvercookedHostVecEnv(ns, num_envs=4)


def test_product_env_refuses_the_emulation_and_cpu_devices():
    """The CPU emulation of the device code is reachable only through the test-side subclass
    (tests/parity_util.EmuVecEnv): the product class accepts nothing but the CUDA library on a CUDA device, and
    the package itself holds no reference to the emulation."""
    import argparse
    from gym_comm_b200 import _cabi
    from gym_comm_b200.vec_env import OvercookedVecEnv
    from tests.parity_util import emu_library
    ns = argparse.Namespace(level="open-divider_tomato", num_agents=2, max_num_timesteps=100)
    with pytest.raises(RuntimeError, match="no other backend"):
        OvercookedVecEnv(ns, num_envs=4, device="cpu", lib=emu_library())
    with pytest.raises(RuntimeError, match="CUDA device only"):
        OvercookedVecEnv(ns, num_envs=4, device="cpu", lib=_cabi.default_library())
    pkg = os.path.join(ROOT, "gym_comm_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            text = open(os.path.join(pkg, f)).read()
            assert "emu_" not in text and "OC_TEST_EMULATION" not in text and "prefix" not in text, f


def test_header_is_plain_c_and_cxx(tmp_path):
    """include/overcooked_b200.h stands alone: C99 and C++17 translation units that include nothing else compile, and
    a C program can name every entry point with the documented signature (pointer types, no torch / CUDA headers)."""
    import subprocess
    inc = os.path.join(ROOT, "include")
    c = tmp_path / "t.c"
    c.write_text('#include "overcooked_b200.h"\n'
                 "int use(oc_env* e, const int32_t* a, int8_t* o8, float* ts, float* r, uint8_t* d, void* s) {\n"
                 "    oc_config cfg; (void)cfg;\n"
                 "    int (*step)(oc_env*, const int32_t*, float*, float*, double*, uint8_t*, float*, uint32_t, void*) = oc_step; (void)step;\n"
                 "    int (*pack)(oc_env*, const float*, int8_t*, float*, void*) = oc_pack_obs_i8; (void)pack;\n"
                 "    return oc_step_host_i8(e, a, o8, ts, r, 0, d, 0, 0, OC_FLAG_AUTO_RESET, s) + oc_reset_host_i8(e, 0, 0, o8, ts, s);\n"
                 "}\n")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I" + inc, "-c", str(c), "-o", str(tmp_path / "t.o")])
    cxx = tmp_path / "t.cpp"
    cxx.write_text('#include "overcooked_b200.h"\nint v() { return OC_ABI_VERSION + (int)sizeof(oc_config); }\n')
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-I" + inc, "-c", str(cxx), "-o", str(tmp_path / "t2.o")])
