"""Host logic: the product's level compiler against the (independent) oracle parser / BFS for all 19
shipped levels, plus config-surface behaviour (defaults for keys the reference forgets, errors)."""
import json

import numpy as np
import pytest

from gym_comm_b200 import levels_data
from gym_comm_b200.arglist import create_arglist, namespace_from_dict
from gym_comm_b200.level_compiler import LevelError, compile_level, parse_subtask
from oracle.spec_model import SpecEnv, parse_subtask as oracle_parse


@pytest.mark.parametrize("name", sorted(levels_data.LEVELS))
def test_compiled_tables_match_oracle(name):
    for n in (2, 4) if "small" not in name and "superwide" not in name else (2,):
        lv = compile_level(name, n)
        o = SpecEnv.__new__(SpecEnv)
        o.n = n
        o._parse_level(levels_data.LEVELS[name])
        assert (lv.width, lv.height, lv.max_path) == (o.W, o.H, o.M)
        assert [lv.xy(c) for c in lv.starts] == o.starts
        assert [lv.xy(c) for c in lv.counters()] == o.counters
        fixed = [(b, lv.xy(c)) for b, c in zip(lv.object_contents, lv.object_cell) if c >= 0]
        assert fixed == o.init_objs
        assert [b for b, c in zip(lv.object_contents, lv.object_cell) if c < 0] == o.random_reps
        assert lv.items == o.items
        pd = lv.path_dist.reshape(lv.ncell, lv.ncell)
        for a in range(lv.ncell):
            for b in range(lv.ncell):
                assert pd[a, b] == o.pd(lv.xy(a), lv.xy(b)), (name, a, b)
        for s, k, g, a0 in zip(lv.subtasks, lv.subtask_kind, lv.subtask_goal, lv.subtask_arg0):
            kind, c, ch, arg0 = oracle_parse(s)
            assert {"Chop": 0, "Merge": 1, "Deliver": 2}[kind] == k and g == (c | ch << 4) and a0 == arg0


def test_subtask_signatures():
    assert parse_subtask("Chop(Tomato)") == (0, 1 | 1 << 4, 1)
    assert parse_subtask("Merge(Tomato, Lettuce-Plate)") == (1, 11 | 3 << 4, 1)
    assert parse_subtask("Deliver(Plate-Tomato)") == (2, 9 | 1 << 4, 9)


def test_arglist_defaults_and_json(tmp_path):
    ns = namespace_from_dict(dict(level="open-divider_tomato", num_agents=2))
    assert ns.num_communication == 10 and ns.fow_radius == 2 and ns.max_num_timesteps == 100
    assert ns.ego_config == dict(CAN_MOVE=True, ALLERGIC=False, BLIND=False) == ns.partner_config
    cfg = dict(level="random-salad-superwide", num_agents=2, max_num_timesteps=900, communication_on=True,
               num_communication=100, ego_config=dict(ALLERGIC=True), hyperparams=dict(n_steps=5000))
    p = tmp_path / "env_args.json"
    p.write_text(json.dumps(cfg))
    ns = create_arglist(str(p))
    assert ns.ego_config == dict(CAN_MOVE=True, ALLERGIC=True, BLIND=False) and ns.num_communication == 100
    assert ns.hyperparams["n_steps"] == 5000
    with pytest.raises(KeyError):
        namespace_from_dict(dict(level="x"))


def test_level_errors():
    with pytest.raises(LevelError):
        compile_level("no-such-level", 2)
    with pytest.raises(LevelError):
        compile_level("random-salad-superwide", 3)            # only two start positions
    two_tomatoes = "-t-t-\n-   -\n-/*p-\n\nSimpleTomato\n\n1 1\n3 1\n"
    with pytest.raises(LevelError):
        compile_level("x", 2, level_text=two_tomatoes)
    with pytest.raises(LevelError):
        compile_level("x", 2, level_text="-t-\n- -\n-*-\n\nNoSuchRecipe\n\n1 1\n1 1\n")


def test_second_line_of_random_objects_is_refused():
    """The reference draws every phase-4 line with its own `occupied` set (overcooked_environment.py:157-166); the
    device draws all random objects without replacement, so a second line would diverge -- refused by name."""
    import pytest
    from gym_comm_b200.level_compiler import LevelError, compile_level
    text = " --- \n-   -\n-   *\n -/- \n\nSimpleTomato\n\n1 1\n2 1\n\npt\np\n"
    with pytest.raises(LevelError, match="phase-4"):
        compile_level("custom", 2, level_text=text)
    ok = compile_level("custom", 2, level_text=text.replace("\np\n", "\n"))
    assert ok.num_random == 2
