"""The reference's OWN PantheonRL code (`pantheonrl/common/multiagentenv.py`: `MultiAgentEnv.step` / `reset`,
partner bookkeeping) drives the drop-in (`gym_comm_b200.compat`, CPU emulation backend) and the live reference env
side by side with the same scripted partner: ego observations, rewards, dones, infos and everything the partner
is shown / told are identical.  Needs /root/reference (build container only)."""
import random

import numpy as np
import pytest

from oracle import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="needs /root/reference")

KEYS = ref_harness.LiveReference.OBS_KEYS


class ScriptedPartner:
    """Duck-typed pantheonrl Agent: a fixed action table; records what it is shown and told."""

    def __init__(self, seed, C):
        self.rng, self.C, self.seen, self.told = random.Random(seed), C, [], []

    def get_action(self, obs, record=True):
        self.seen.append({k: np.asarray(obs.obs[k], dtype=np.float64).copy() for k in KEYS})
        return (self.rng.randrange(4), self.rng.randrange(self.C))

    def update(self, reward, done):
        self.told.append((float(reward), bool(done)))


@pytest.mark.parametrize("level,T,C", [("open-divider_tomato", 9, 5), ("open-divider_salad", 14, 3)])
def test_reference_multiagentenv_drives_the_drop_in(level, T, C):
    from gym_comm_b200 import compat
    from tests.parity_util import EmuMultiEnv
    ns = ref_harness.make_namespace(level, max_num_timesteps=T, num_communication=C)
    live = ref_harness.LiveReference(ns, py_random_seed=3)            # installs the import stubs, builds the reference env
    ref_env = live.wrapper
    Env = compat.gym_env_class(EmuMultiEnv)
    from pantheonrl.common.multiagentenv import SimultaneousEnv
    assert issubclass(Env, SimultaneousEnv)
    ours = Env(ns, device="cpu", level_text=live.level_text(), subtasks=live.subtask_strings())
    p_ref, p_ours = ScriptedPartner(1, C), ScriptedPartner(1, C)
    ref_env.add_partner_agent(p_ref)
    ours.add_partner_agent(p_ours)
    rng = random.Random(0)

    def same(a, b):
        return all(np.array_equal(np.asarray(a[k], dtype=np.float64), np.asarray(b[k], dtype=np.float64)) for k in KEYS)

    for episode in range(3):
        with ref_harness._in_cwd(live.cwd), ref_harness.quiet():
            o_ref = ref_env.reset()
        o = ours.reset()
        assert same(o, o_ref), episode
        done = False
        while not done:
            act = (rng.randrange(4), rng.randrange(C))
            with ref_harness._in_cwd(live.cwd), ref_harness.quiet():
                o_ref, r_ref, d_ref, i_ref = ref_env.step(act)
            o, r, done, info = ours.step(act)
            assert same(o, o_ref) and r == r_ref and done == d_ref and info == i_ref
    assert len(p_ours.seen) == len(p_ref.seen) == 3 * T and p_ours.told == p_ref.told
    assert all(same(a, b) for a, b in zip(p_ours.seen, p_ref.seen))
    ours.close()
