"""TEST INFRASTRUCTURE ONLY -- records golden traces from the LIVE reference.

Run in the build container (the reference is not present on the GPU box):

    PYTHONHASHSEED=0 python -m oracle.record_golden            # writes tests/golden/*.npz

Each trace holds the config, the level text, the reference's subtask table, the
action/message sequence, the object placements of every reset, and per step the
reference's returned reward (f64), done, the flat observation of every observer
(key-sorted order) and a canonical state dump.  While recording,
the Python restatement (``oracle/spec_model.py``) runs in lock-step -- it only
supplies the state the goal-chasing action source looks at, and the recorder
asserts it never diverges, so a recorded trace is also a passed parity run.

Reference entry points driven: ``OvercookedMultiEnv.multi_step`` /
``multi_reset`` (gym_comm/envs/overcooked_env.py:207-297) for 2 agents;
``OvercookedEnvironment.step`` + ``get_observation2`` for 3-4 agents.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

from .drivers import GoalChaser
from .level_fuzz import random_level
from .ref_harness import LiveReference, hashseed_is_canonical, make_namespace
from .spec_model import BIT, SpecEnv

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

ALLERGIC_EGO = dict(CAN_MOVE=False, ALLERGIC=True, BLIND=False)      # spread/env_args20on_allergic.json
BLIND_PARTNER = dict(CAN_MOVE=True, ALLERGIC=False, BLIND=True)

CUSTOM_LEVEL = """-p/--*--
-      -
t  --  l
-      /
- -  - -
*      -
---p----

Salad

1 1
6 1
3 3
4 5
"""

ONION_LEVEL = """--/--*-
t     l
-     o
/     p
---p-p-

OnionSalad

1 1
5 1
3 2
"""

SCENARIOS = [
    # name, level, steps, kwargs
    ("tomato_a9_script", "open-divider_tomato", 0, dict(num_communication=5, max_num_timesteps=500, script="a9")),
    ("tomato_chaser", "open-divider_tomato", 700, dict(max_num_timesteps=100)),
    ("partial_salad_2a", "partial-divider_salad", 900, dict(max_num_timesteps=200)),
    ("full_salad_3a", "full-divider_salad", 700, dict(num_agents=3, max_num_timesteps=200)),
    ("partial_salad_3a", "partial-divider_salad", 700, dict(num_agents=3, max_num_timesteps=200)),
    ("open_salad_4a", "open-divider_salad", 500, dict(num_agents=4, max_num_timesteps=200)),
    ("open_tl", "open-divider_tl", 900, dict(max_num_timesteps=200)),
    ("cramped_allergic", "random-open-divider_salad_small_cramped", 900,
     dict(max_num_timesteps=150, num_communication=8, fow_radius=10,
          ego_config=ALLERGIC_EGO, partner_config=BLIND_PARTNER)),
    ("superwide_c100", "random-salad-superwide", 350, dict(max_num_timesteps=120, num_communication=100)),
    ("small_wide_commoff", "random-open-divider_salad_small_wide", 500,
     dict(max_num_timesteps=150, communication_on=False)),
    ("open_salad_egoled_allergic_partner", "open-divider_salad", 500,
     dict(max_num_timesteps=150, ego_led=True, partner_config=dict(ALLERGIC=True))),
    ("random_small", "random-open-divider_salad_small", 500, dict(max_num_timesteps=120, fow_radius=1)),
    ("random_tomato_fow0", "random-open-divider_tomato", 500, dict(max_num_timesteps=120, fow_radius=0)),
    ("wide_big_blind_ego", "random-open-divider_salad_small_wide_big", 400,
     dict(max_num_timesteps=120, ego_config=dict(BLIND=True))),
    ("custom_level", "custom-two-deliveries", 700, dict(max_num_timesteps=200, level_text=CUSTOM_LEVEL)),
    # rows whose width is not a multiple of 4 floats (3 observers x 49 features)
    ("open_tl_3a_c5", "open-divider_tl", 500, dict(num_agents=3, max_num_timesteps=150, num_communication=5, fow_radius=3)),
    # random kitchens (oracle/level_fuzz.py): geometry the shipped levels never exercise
    ("fuzz_kitchen_a", "fuzz-a", 400, dict(max_num_timesteps=90, num_communication=4, fow_radius=1,
                                            level_text=random_level(1003))),
    ("fuzz_kitchen_b", "fuzz-b", 400, dict(num_agents=3, max_num_timesteps=80, num_communication=6, fow_radius=2,
                                            level_text=random_level(1008, 3))),
    # three Foods, 29 subtasks, 6 shaping item pairs
    ("onion_salad_2a", "onion-salad", 900, dict(max_num_timesteps=250, num_communication=6, level_text=ONION_LEVEL)),
    ("fuzz_kitchen_c", "fuzz-c", 400, dict(max_num_timesteps=100, num_communication=7, fow_radius=3, ego_led=True,
                                            level_text=random_level(1011))),
]

A9_PARTNER = [3, 1, 0, 2, 2, 2, 2, 2, 0, 0, 0, 3, 3, 3, 3, 3, 1, 1, 2, 2, 2, 2, 2]


def _placements(ref, reps):
    if not reps:
        return None
    by_bits = {BIT[nm]: loc for nm, loc in ref.object_placements()}
    return [by_bits[b] for b in reps]


def _state_arrays(spec, nobj):
    objs = np.full((nobj, 5), -1, dtype=np.int16)
    for i, o in enumerate(spec.ordered()):
        objs[i] = (o.contents, o.chopped, o.loc[0], o.loc[1], int(o.held))
    agents = np.array(spec.agents, dtype=np.int16)
    hold = np.array([(-1 if h is None else h.contents | (h.chopped << 4)) for h in spec.hold], dtype=np.int16)
    return objs, agents, hold


def record(name, level, steps, kw, seed):
    kw = dict(kw)
    script = kw.pop("script", None)
    level_text = kw.pop("level_text", None)
    n = kw.get("num_agents", 2)
    ns = make_namespace(level, **kw)
    ref = LiveReference(ns, py_random_seed=seed, level_text=level_text)
    text = ref.level_text()
    subtasks = ref.subtask_strings()
    probe = SpecEnv.__new__(SpecEnv)
    probe.n = n
    probe._parse_level(text)
    reps = probe.random_reps
    pl0 = _placements(ref, reps)
    spec = SpecEnv(text, subtasks, num_agents=n, max_num_timesteps=ns.max_num_timesteps,
                   communication_on=ns.communication_on, num_communication=ns.num_communication,
                   ego_led=ns.ego_led, fow_radius=ns.fow_radius, ego_config=ns.ego_config,
                   partner_config=ns.partner_config, placements=pl0)
    nobj = len(spec.objs)
    chaser = GoalChaser(spec, seed=seed)
    if script == "a9":
        acts = [([1, A9_PARTNER[i]], [i % 5, (2 * i) % 5]) for i in range(len(A9_PARTNER))]
        steps = len(acts)

    def check_and_obs():
        assert spec.state_tuple() == ref.state_tuple(), (name, spec.state_tuple(), ref.state_tuple())
        rows = []
        for k in range(n):
            a = ref.flat_obs(k)
            assert list(a) == spec.flat_obs(k), (name, k)
            rows.append(a)
        return np.stack(rows)

    out = dict(navs=[], comms=[], reward=[], done=[], obs=[], reset_obs=[check_and_obs()],
               placements=[pl0 if pl0 is not None else []], objs=[], agents=[], hold=[],
               completed=[], counts=[], t=[])
    events = dict(sparse=0, dones=0)
    for i in range(steps):
        navs, comms = acts[i] if script else chaser.act()
        r, d = ref.step(navs, comms)
        r2, d2, sp = spec.step(navs, comms)
        assert r == r2 and d == d2, (name, i, r, r2, d, d2)
        out["navs"].append(navs)
        out["comms"].append(comms)
        out["reward"].append(r)
        out["done"].append(d)
        out["obs"].append(check_and_obs())
        o, a, h = _state_arrays(spec, nobj)
        out["objs"].append(o)
        out["agents"].append(a)
        out["hold"].append(h)
        out["completed"].append(list(spec.completed))
        out["counts"].append(list(spec.count))
        out["t"].append(spec.t)
        events["sparse"] += int(sp != 0)
        if d:
            events["dones"] += 1
            ref.reset()
            pl = _placements(ref, reps)
            spec.reset(pl)
            chaser.on_reset()
            out["reset_obs"].append(check_and_obs())
            out["placements"].append(pl if pl is not None else [])
    meta = dict(name=name, level=level, level_text=text, subtasks=subtasks, num_agents=n,
                max_num_timesteps=ns.max_num_timesteps, communication_on=ns.communication_on,
                num_communication=ns.num_communication, ego_led=ns.ego_led, fow_radius=ns.fow_radius,
                ego_config=ns.ego_config, partner_config=ns.partner_config, seed=seed,
                random_reps=reps, events=events,
                source="live reference, PYTHONHASHSEED=0, oracle/record_golden.py")
    arrays = dict(
        navs=np.array(out["navs"], dtype=np.int8), comms=np.array(out["comms"], dtype=np.int16),
        reward=np.array(out["reward"], dtype=np.float64), done=np.array(out["done"], dtype=np.bool_),
        obs=np.stack(out["obs"]).astype(np.float64), reset_obs=np.stack(out["reset_obs"]).astype(np.float64),
        placements=np.array(out["placements"], dtype=np.int16).reshape(len(out["placements"]), len(reps), 2),
        objs=np.stack(out["objs"]), agents=np.stack(out["agents"]), hold=np.stack(out["hold"]),
        completed=np.array(out["completed"], dtype=np.int8), counts=np.array(out["counts"], dtype=np.int8),
        t=np.array(out["t"], dtype=np.int32), meta=np.array(json.dumps(meta)))
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-38s steps=%4d sparse_events=%3d dones=%2d  %6.1f KB" %
          (name, steps, events["sparse"], events["dones"], os.path.getsize(path) / 1024), file=sys.stderr)


def main():
    assert hashseed_is_canonical(), "run with PYTHONHASHSEED=0 (subtask order depends on it)"
    only = sys.argv[1:]
    for idx, (name, level, steps, kw) in enumerate(SCENARIOS):
        if only and name not in only:
            continue
        record(name, level, steps, kw, seed=1000 + idx)


if __name__ == "__main__":
    main()
