"""`OvercookedVecEnv.render` reproduces `str(OvercookedEnvironment)` of the live reference.  Runs
the device code through the CPU emulation (tests/emu), so it needs no GPU; skipped without
/root/reference.  Subtask order is irrelevant for the picture, so any hash seed works."""
import pytest
import torch

from oracle import ref_harness
from oracle.drivers import GoalChaser
from oracle.spec_model import SpecEnv
from gym_comm_b200.vec_env import OvercookedVecEnv
from tests.parity_util import EmuVecEnv, emu_library

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="needs /root/reference")


def test_render_matches_reference_display():
    ns = ref_harness.make_namespace("partial-divider_salad", max_num_timesteps=400)
    ref = ref_harness.LiveReference(ns)
    spec = SpecEnv(ref.level_text(), ref.subtask_strings(), max_num_timesteps=400)
    env = EmuVecEnv(ns, num_envs=1, device="cpu", auto_reset=False, lib=emu_library(),
                           subtasks=ref.subtask_strings())
    chaser = GoalChaser(spec, seed=4, p_random=0.1)
    # (right after reset() the reference's display buffer is empty until the first step, :188)
    seen_merged = False
    for i in range(300):
        navs, comms = chaser.act()
        ref.step(navs, comms)
        _, done, _ = spec.step(navs, comms)
        env.step(torch.tensor([[[navs[0], comms[0]], [navs[1], comms[1]]]], dtype=torch.int32))
        assert env.render(0) == str(ref.base), (i, env.render(0), str(ref.base))
        seen_merged |= any(bin(o.contents).count("1") > 1 for o in spec.objs if o.alive)
        if done:
            break
    assert seen_merged
