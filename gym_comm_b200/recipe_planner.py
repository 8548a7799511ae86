"""Subtask derivation: the reference's STRIPS planner, restated for compile time (host side, pure Python).

Reference: ``OvercookedEnvironment.run_recipes``  gym_cooking/envs/overcooked_environment.py:452-459
           ``STRIPSWorld.__init__ / generate_graph / get_subtasks``  gym_cooking/recipe_planner/stripsworld.py:12-79
           ``Recipe.add_ingredient / add_goal / add_merge_actions``  gym_cooking/recipe_planner/recipe.py:5-66
           actions and predicates  gym_cooking/recipe_planner/utils.py:16-165

What the reference computes, per recipe: a breadth-first search over multisets of predicates from the initial state
(``None`` + one ``Fresh(X)`` per world object that contains X) until the first state that holds
``Delivered(<all ingredients>-Plate)``; the recipe's subtasks are the UNION of the actions on ALL shortest paths to
that state.  ``all_subtasks`` is the concatenation over the recipe list.

Two things in the reference depend on Python's string-hash seed, and this module makes both explicit instead of
imitating CPython's set internals:

* the ORDER of a recipe's subtasks is the iteration order of a Python ``set`` of actions.  `derive_subtasks` returns
  them in a canonical order (sorted strings); ``levels_data.SUBTASKS`` holds the reference's own order at
  ``PYTHONHASHSEED=0`` for every list of its four recipes, and `order_like` re-orders a derived list to match it.
* two actions can label the SAME edge (``Merge(Tomato, Lettuce)`` and ``Merge(Lettuce, Tomato)`` both lead from
  {Chopped(Tomato), Chopped(Lettuce)} to {Merged(Lettuce-Tomato)}); networkx keeps whichever ``graph.add_edge`` came
  last (stripsworld.py:44), i.e. one of them, by hash order.  `derive_subtasks` reports such a pair as one subtask
  with the operands sorted, and lists the alternatives in ``Derived.either``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from itertools import combinations
from typing import Dict, FrozenSet, Iterable, List, Sequence, Tuple

# recipe.py:68-97 -- every ingredient is FRESH_CHOPPED (core.py FoodSequence), contents sorted by name (recipe.py:29)
RECIPE_CONTENTS = {
    "SimpleTomato": ("Tomato",),
    "SimpleLettuce": ("Lettuce",),
    "Salad": ("Lettuce", "Tomato"),
    "OnionSalad": ("Lettuce", "Onion", "Tomato"),
}
OBJECT_NAMES = ("Plate", "Tomato", "Lettuce", "Onion")      # stripsworld.py:21

NONE = "None"


def _join(names: Iterable[str]) -> str:
    return "-".join(sorted(names))


@dataclass(frozen=True)
class Action:
    """name(args) with its precondition / postcondition predicate lists (utils.py:57-165); predicates are strings."""
    name: str
    args: Tuple[str, ...]
    pre: Tuple[str, ...]
    post: Tuple[str, ...]

    def __str__(self) -> str:
        return "%s(%s)" % (self.name, ", ".join(self.args))


def _get(obj):                                                # utils.py:101-109
    return Action("Get", (obj,), (NONE,), ("Fresh(%s)" % obj, NONE))


def _chop(obj):                                               # utils.py:116-124
    return Action("Chop", (obj,), ("Fresh(%s)" % obj,), ("Chopped(%s)" % obj,))


def _merge(a, b, pre=None):                                   # utils.py:131-141
    if pre is None:
        pre = ("Chopped(%s)" % a, "Merged(%s)" % b)
    return Action("Merge", (a, b), tuple(pre), ("Merged(%s)" % _join(a.split("-") + b.split("-")),))


def _deliver(obj):                                            # utils.py:147-152
    return Action("Deliver", (obj,), ("Merged(%s)" % obj,), ("Delivered(%s)" % obj,))


def recipe_actions(recipe: str) -> Tuple[Dict[Tuple[str, Tuple[str, ...]], Action], str]:
    """-> (the recipe's action set keyed by (name, args) -- Action.__eq__ / __hash__ ignore pre / post, so the FIRST
    action added under a key stays, utils.py:78-83 --, the goal predicate)."""
    if recipe not in RECIPE_CONTENTS:
        raise ValueError("unknown recipe %r" % (recipe,))
    names = RECIPE_CONTENTS[recipe]
    acts: Dict[Tuple[str, Tuple[str, ...]], Action] = {}

    def add(a: Action):
        acts.setdefault((a.name, a.args), a)

    add(_get("Plate"))                                                               # recipe.py:10
    for n in names:                                                                  # add_ingredient, recipe.py:15-25
        add(_get(n))
        add(_chop(n))
        add(_merge(n, "Plate", ("Chopped(%s)" % n, "Fresh(Plate)")))
    full_plate = _join(names + ("Plate",))                                           # add_goal, recipe.py:27-33
    add(_deliver(full_plate))
    for i in range(2, len(names) + 1):                                               # add_merge_actions, recipe.py:35-66
        for combo in combinations(names, i):
            add(_merge(_join(combo), "Plate", ("Merged(%s)" % _join(combo), "Fresh(Plate)")))
            for item in combo:
                rem = [c for c in combo if c != item]
                rem_str, plate_str, rem_plate = _join(rem), _join([item, "Plate"]), _join(rem + ["Plate"])
                if len(rem) == 1:
                    add(_merge(item, rem_str, ("Chopped(%s)" % item, "Chopped(%s)" % rem_str)))
                    add(_merge(rem_str, plate_str))
                    add(_merge(item, rem_plate))
                else:
                    add(_merge(item, rem_str))
                    add(_merge(plate_str, rem_str, ("Merged(%s)" % plate_str, "Merged(%s)" % rem_str)))
                    add(_merge(item, rem_plate))
    return acts, "Delivered(%s)" % full_plate


State = Tuple[str, ...]          # sorted multiset of predicate strings (STRIPSState.__eq__ / __hash__, utils.py:164-171)


def _apply(state: State, a: Action):
    """Action.is_valid_in + get_next_from (utils.py:85-99): every precondition removes ONE instance."""
    s = list(state)
    for pre in a.pre:
        try:
            s.remove(pre)
        except ValueError:
            return None
    s.extend(a.post)
    return tuple(sorted(s))


@dataclass
class Derived:
    subtasks: List[str]                                  # canonical order (sorted), one entry per edge-label class
    either: Dict[str, Tuple[str, ...]] = field(default_factory=dict)   # entry -> the labels the reference may show instead
    depth: int = 0


def derive_recipe_subtasks(recipe: str, world_object_contents: Sequence[Sequence[str]], max_path_length: int = 14) -> Derived:
    """The subtasks of ONE recipe in a world whose objects hold `world_object_contents` (a list of name lists, one per
    object: a plated tomato is ["Tomato", "Plate"])."""
    acts, goal = recipe_actions(recipe)
    init = [NONE]
    for contents in world_object_contents:                                           # stripsworld.py:19-23
        for n in OBJECT_NAMES:
            if n in contents:
                init.append("Fresh(%s)" % n)
    init = tuple(sorted(init))
    depth = {init: 0}
    edges: Dict[Tuple[State, State], List[Action]] = {}
    level = [init]
    goal_states: List[State] = []
    d = 0
    while level and not goal_states and d < max_path_length:
        d += 1
        nxt = []
        for s in level:
            for a in acts.values():
                t = _apply(s, a)
                if t is None:
                    continue
                if t not in depth:
                    depth[t] = d
                    nxt.append(t)
                if depth[t] == d:                       # a forward edge of the BFS layering: candidates for shortest paths
                    edges.setdefault((s, t), []).append(a)
                    if goal in t and t not in goal_states:
                        goal_states.append(t)
        level = nxt
    if not goal_states:
        raise ValueError("recipe %s: no plan within max_num_subtasks = %d steps (the reference exits, stripsworld.py:55-57)"
                         % (recipe, max_path_length))
    # every path of `d` steps to a goal state spends the same ingredients, so there is ONE goal state at the first depth
    # that has any; should a custom recipe ever produce several, the reference would take whichever its set order
    # visits first -- refuse rather than guess
    if len(goal_states) != 1:
        raise ValueError("recipe %s: %d distinct goal states at depth %d -- the reference's choice depends on its hash seed"
                         % (recipe, len(goal_states), d))
    # backward sweep: edges that lie on a shortest path initial -> goal
    on_path = {goal_states[0]}
    by_target: Dict[State, List[Tuple[State, List[Action]]]] = {}
    for (s, t), al in edges.items():
        by_target.setdefault(t, []).append((s, al))
    frontier = [goal_states[0]]
    classes: List[FrozenSet[str]] = []
    while frontier:
        new = []
        for t in frontier:
            for s, al in by_target.get(t, ()):
                classes.append(frozenset(str(a) for a in al))
                if s not in on_path:
                    on_path.add(s)
                    new.append(s)
        frontier = new
    out = Derived(subtasks=[], depth=d)
    seen = set()
    for c in classes:
        if c in seen:
            continue
        seen.add(c)
        rep = canonical_label(sorted(c)[0])
        if len(c) > 1:
            out.either[rep] = tuple(sorted(c))
        out.subtasks.append(rep)
    # a label that is alone on one edge and one of a pair on another cannot happen (a pair shares source and target)
    out.subtasks = sorted(set(out.subtasks))
    return out


def canonical_label(label: str) -> str:
    """Merge(X, Y) of two bare chopped foods is symmetric (both operand orders are actions of the recipe and label the
    same edge): canonical form = operands sorted.  Everything else is returned unchanged."""
    if label.startswith("Merge("):
        a, b = label[6:-1].split(", ")
        if "-" not in a and "-" not in b and a != "Plate" and b != "Plate":
            a, b = sorted((a, b))
            return "Merge(%s, %s)" % (a, b)
    return label


def derive_subtasks(recipes: Sequence[str], world_object_contents: Sequence[Sequence[str]],
                    max_path_length: int = 14) -> List[str]:
    """``all_subtasks`` (overcooked_environment.py:457) with every recipe's subtasks in canonical order."""
    out: List[str] = []
    for r in recipes:
        out += derive_recipe_subtasks(r, world_object_contents, max_path_length).subtasks
    return out


def order_like(derived: Sequence[str], recorded: Sequence[str]) -> List[str]:
    """`recorded` (the reference's list at PYTHONHASHSEED=0) if it is `derived` up to order and symmetric-merge operand
    order; raises ValueError otherwise.  This is the cross-check between the planner and ``levels_data.SUBTASKS``."""
    a = sorted(canonical_label(s) for s in derived)
    b = sorted(canonical_label(s) for s in recorded)
    if a != b:
        raise ValueError("derived subtasks %r differ from the recorded table %r" % (a, b))
    return list(recorded)
