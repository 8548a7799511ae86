"""Lock-step of the Python restatement against the LIVE reference (build container only).

Skipped when /root/reference is absent (GPU box) or PYTHONHASHSEED != 0 (the reference's
subtask order depends on it).  The committed golden traces are the portable
form of this check."""
import random

import pytest

from oracle import ref_harness
from oracle.drivers import GoalChaser
from oracle.spec_model import BIT, SpecEnv

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="needs /root/reference")

if ref_harness.reference_available() and not ref_harness.hashseed_is_canonical():
    # the hash seed is fixed at interpreter start: re-run this module in a child with PYTHONHASHSEED=0
    import os
    import subprocess
    import sys

    def test_lockstep_in_child_with_canonical_hashseed():
        env = dict(os.environ, PYTHONHASHSEED="0")
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", os.path.abspath(__file__)], env=env, cwd=root,
                           capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
        assert "4 passed" in r.stdout, r.stdout[-500:]

    CASES_ENABLED = False
else:
    CASES_ENABLED = True

CASES = [
    ("open-divider_tomato", 250, dict(max_num_timesteps=60)),
    ("partial-divider_salad", 250, dict(max_num_timesteps=90, num_agents=3)),
    ("random-salad-superwide", 200, dict(max_num_timesteps=70, num_communication=100)),
    ("random-open-divider_salad_small_cramped", 250,
     dict(max_num_timesteps=80, num_communication=8, fow_radius=10,
          ego_config=dict(CAN_MOVE=False, ALLERGIC=True, BLIND=False),
          partner_config=dict(CAN_MOVE=True, ALLERGIC=False, BLIND=True))),
]


@pytest.mark.skipif(not CASES_ENABLED, reason="runs in the PYTHONHASHSEED=0 child")
@pytest.mark.parametrize("level,steps,kw", CASES)
def test_lockstep(level, steps, kw):
    n = kw.get("num_agents", 2)
    ns = ref_harness.make_namespace(level, **kw)
    ref = ref_harness.LiveReference(ns, py_random_seed=7)
    probe = SpecEnv.__new__(SpecEnv)
    probe.n = n
    probe._parse_level(ref.level_text())

    def placements():
        if not probe.random_reps:
            return None
        by = {BIT[nm]: loc for nm, loc in ref.object_placements()}
        return [by[b] for b in probe.random_reps]

    spec = SpecEnv(ref.level_text(), ref.subtask_strings(), num_agents=n,
                   max_num_timesteps=ns.max_num_timesteps, communication_on=ns.communication_on,
                   num_communication=ns.num_communication, ego_led=ns.ego_led, fow_radius=ns.fow_radius,
                   ego_config=ns.ego_config, partner_config=ns.partner_config, placements=placements())
    chaser = GoalChaser(spec, seed=random.Random(level).randrange(1 << 30))
    for i in range(steps):
        navs, comms = chaser.act()
        r1, d1 = ref.step(navs, comms)
        r2, d2, _ = spec.step(navs, comms)
        assert (r1, d1) == (r2, d2), i
        assert spec.state_tuple() == ref.state_tuple(), i
        for k in range(n):
            assert list(ref.flat_obs(k)) == spec.flat_obs(k), (i, k)
        if d1:
            ref.reset()
            spec.reset(placements())
            chaser.on_reset()
            for k in range(n):
                assert list(ref.flat_obs(k)) == spec.flat_obs(k)
