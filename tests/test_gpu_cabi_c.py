"""GPU: the C ABI driven from a plain C program (tests/cabi_smoke.c) -- no Python, no torch in the
loop: hand-built oc_config, oc_create with library-side path tabulation, oc_step / oc_rollout /
oc_reset, error codes."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_program_drives_the_abi(tmp_path):
    import __graft_entry__ as ge
    ge.build_cuda()
    exe = str(tmp_path / "cabi_smoke")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    libdir = os.path.join(ROOT, "gym_comm_b200")
    subprocess.check_call([nvcc, "-x", "c", os.path.join(ROOT, "tests", "cabi_smoke.c"), "-o", exe,
                           "-L" + libdir, "-loc_b200", "-Xlinker", "-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "cabi_smoke ok" in out.stdout
