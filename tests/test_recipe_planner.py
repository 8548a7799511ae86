"""The subtask derivation (gym_comm_b200/recipe_planner.py = the reference's STRIPSWorld.get_subtasks,
recipe_planner/stripsworld.py:61-79) against the tables recorded from the live reference (levels_data.SUBTASKS,
tools/gen_levels_data.py): the same subtasks for every list of the four recipes, up to order and the operand order of
symmetric merges (both hash-seed dependent in the reference)."""
import itertools

import pytest

from gym_comm_b200 import levels_data
from gym_comm_b200.level_compiler import LevelError, compile_level, derive_level_subtasks
from gym_comm_b200.recipe_planner import (canonical_label, derive_recipe_subtasks, derive_subtasks, order_like,
                                          recipe_actions)

NAMES = ("SimpleTomato", "SimpleLettuce", "Salad", "OnionSalad")
WORLD = [["Tomato"], ["Lettuce"], ["Onion"], ["Plate"], ["Plate"]]


def test_planner_reproduces_every_recorded_table():
    n = 0
    for k in range(1, 5):
        for recipes in itertools.permutations(NAMES, k):
            table = levels_data.SUBTASKS[recipes]
            derived = derive_subtasks(recipes, WORLD)
            assert order_like(derived, table) == table
            assert len(derived) == len(table)
            n += 1
    assert n == 64


def test_plan_depth_and_symmetric_merges():
    t = derive_recipe_subtasks("SimpleTomato", WORLD)
    assert t.depth == 3 and t.subtasks == ["Chop(Tomato)", "Deliver(Plate-Tomato)", "Merge(Tomato, Plate)"] and not t.either
    s = derive_recipe_subtasks("Salad", WORLD)
    assert s.depth == 5 and len(s.subtasks) == 9
    assert s.either == {"Merge(Lettuce, Tomato)": ("Merge(Lettuce, Tomato)", "Merge(Tomato, Lettuce)")}
    o = derive_recipe_subtasks("OnionSalad", WORLD)
    assert o.depth == 7 and len(o.subtasks) == 29 and len(o.either) == 3
    assert canonical_label("Merge(Tomato, Lettuce)") == "Merge(Lettuce, Tomato)"
    assert canonical_label("Merge(Tomato, Lettuce-Plate)") == "Merge(Tomato, Lettuce-Plate)"


def test_world_contents_shape_the_plan():
    # extra objects do not matter; a plated object counts for each of its contents (stripsworld.py:19-23)
    assert derive_subtasks(["Salad"], WORLD) == derive_subtasks(["Salad"], [["Tomato", "Plate"], ["Lettuce"], ["Plate"]])
    # a missing ingredient has to be fetched: one more step, a Get subtask -- which the env has no goal objects for
    d = derive_recipe_subtasks("Salad", [["Tomato"], ["Plate"]])
    assert d.depth == 6 and "Get(Lettuce)" in d.subtasks
    with pytest.raises(LevelError, match="Lettuce"):
        compile_level("custom", 2, level_text="--/--*-\nt     -\n-     -\n-     p\n-------\n\nSalad\n\n1 1\n5 1\n")
    with pytest.raises(ValueError, match="no plan"):
        derive_recipe_subtasks("OnionSalad", WORLD, max_path_length=5)


def test_action_sets():
    acts, goal = recipe_actions("Salad")
    assert goal == "Delivered(Lettuce-Plate-Tomato)"
    names = sorted(str(a) for a in acts.values())
    assert "Merge(Tomato, Lettuce)" in names and "Merge(Lettuce, Tomato)" in names and "Get(Plate)" in names
    assert len(names) == 3 + 2 * 2 + 1 + 5      # Get x3, Chop + Merge-with-Plate per food, Deliver, 5 more merges (recipe.py:35-66)
    with pytest.raises(ValueError):
        recipe_actions("Soup")


def test_compile_level_uses_the_planner_in_table_order():
    for name in ("open-divider_tomato", "partial-divider_salad", "full-divider_onionsalad"):
        if name not in levels_data.LEVELS:
            continue
        lv = compile_level(name, 2)
        assert lv.subtasks == levels_data.SUBTASKS[lv.recipes]
    assert derive_level_subtasks(("Salad", "Salad"), [1, 2, 8, 8]) == levels_data.SUBTASKS[("Salad",)] * 2
