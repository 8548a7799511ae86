"""Level compiler: level text + recipe names -> the static tables the CUDA kernels consume.

Mirrors what the reference recomputes on EVERY ``reset()`` (125-320 ms) and does it
once per configuration:

* ``OvercookedEnvironment.load_level``      gym_cooking/envs/overcooked_environment.py:100-178
* ``run_recipes`` / ``STRIPSWorld.get_subtasks``  :452-459, recipe_planner/stripsworld.py:61-79
  (derived by ``recipe_planner.derive_recipe_subtasks`` from the recipe list and the objects of the level; the ORDER
  is PYTHONHASHSEED-dependent in the reference and is taken from ``levels_data.SUBTASKS``, recorded at the canonical
  seed 0, which the derivation must reproduce as a set; pass ``subtasks=`` to override)
* ``get_subtask_obj`` goal templates        gym_cooking/navigation_planner/utils.py:161-209
* ``World.make_reachability_graph`` + ``get_path_distance_between``  gym_cooking/utils/world.py:61-131
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from functools import lru_cache

from . import levels_data, recipe_planner

# encodings shared with include/overcooked_b200.h
TILE_FLOOR, TILE_COUNTER, TILE_CUTBOARD, TILE_DELIVERY = 0, 1, 2, 3
CONTENT_BIT = {"Tomato": 1, "Lettuce": 2, "Onion": 4, "Plate": 8}      # 1 << ObjectChannel (core.py:383-388)
REP_BIT = {"t": 1, "l": 2, "o": 4, "p": 8}                              # core.py:18-26
REP_TILE = {"-": TILE_COUNTER, "/": TILE_CUTBOARD, "*": TILE_DELIVERY}
NAV_ACTIONS = ((0, 1), (0, -1), (-1, 0), (1, 0))                        # world.py:16
SUBTASK_KIND = {"Chop": 0, "Merge": 1, "Deliver": 2}
RECIPE_FOODS = {"SimpleTomato": 1, "SimpleLettuce": 2, "Salad": 3, "OnionSalad": 7}   # recipe.py:68-97
_NAME_ORDER = ("Lettuce", "Onion", "Plate", "Tomato")                   # alphabetical, as Object.name sorts

MAX_OBJECTS, MAX_SUBTASKS, MAX_CELLS, MAX_AGENTS = 6, 32, 128, 4


class LevelError(ValueError):
    pass


@dataclass
class CompiledLevel:
    name: str
    width: int
    height: int
    tiles: np.ndarray                 # uint8 [H*W]
    starts: List[int]                 # start cell per agent, file order
    object_contents: List[int]        # world insertion order (phase 1 reading order, then phase 4)
    object_cell: List[int]            # -1 = random counter at reset
    recipes: Tuple[str, ...]
    subtasks: List[str]
    subtask_kind: List[int] = field(default_factory=list)
    subtask_goal: List[int] = field(default_factory=list)
    subtask_arg0: List[int] = field(default_factory=list)
    items: List[int] = field(default_factory=list)
    max_path: int = 0
    path_dist: Optional[np.ndarray] = None      # uint8 [ncell*ncell]

    @property
    def ncell(self) -> int:
        return self.width * self.height

    @property
    def num_random(self) -> int:
        return sum(1 for c in self.object_cell if c < 0)

    def cell(self, x: int, y: int) -> int:
        return y * self.width + x

    def xy(self, cell: int) -> Tuple[int, int]:
        return cell % self.width, cell // self.width

    def counters(self) -> List[int]:
        return [i for i in range(self.ncell) if self.tiles[i] == TILE_COUNTER]


def parse_subtask(s: str) -> Tuple[int, int, int]:
    """'Merge(Tomato, Lettuce-Plate)' -> (kind, goal signature, arg0 bits).

    Goal signature = contents | chopped << 4 of the template ``get_subtask_obj`` builds: every
    Food of the union in its last (chopped) state (navigation_planner/utils.py:161-209)."""
    kind = s[:s.index("(")]
    args = s[s.index("(") + 1:-1].split(", ")
    if kind not in SUBTASK_KIND:
        raise LevelError("unsupported subtask %r" % s)
    bits = 0
    for a in args:
        for n in a.split("-"):
            bits |= CONTENT_BIT[n]
    arg0 = 0
    for n in args[0].split("-"):
        arg0 |= CONTENT_BIT[n]
    return SUBTASK_KIND[kind], bits | ((bits & 7) << 4), arg0


def path_distance_table(tiles: np.ndarray, width: int, height: int) -> np.ndarray:
    """pd[src, dst] for every cell pair (world.py:114-131): MAX_PATH when src is not a floor tile
    (its node is missing from the reachability graph, the exception is swallowed) or dst cannot
    be reached; BFS over floor cells otherwise; a collidable dst costs 1 + the nearest floor
    neighbour (its graph nodes hang off the adjacent floor nodes, world.py:72-88)."""
    n = width * height
    m = 2 * (width + height) + 1
    pd = np.full((n, n), m, dtype=np.int32)
    floor = [tiles[i] == TILE_FLOOR for i in range(n)]

    def neighbours(c):
        x, y = c % width, c // width
        for dx, dy in NAV_ACTIONS:
            vx, vy = x + dx, y + dy
            if 0 <= vx < width and 0 <= vy < height:
                yield vy * width + vx

    for src in range(n):
        if not floor[src]:
            continue
        dist = {src: 0}
        q = deque([src])
        while q:
            u = q.popleft()
            for v in neighbours(u):
                if floor[v] and v not in dist:
                    dist[v] = dist[u] + 1
                    q.append(v)
        for dst in range(n):
            if floor[dst]:
                best = dist.get(dst, m)
            else:
                best = min([dist[v] + 1 for v in neighbours(dst) if floor[v] and v in dist], default=m)
            pd[src, dst] = min(best, m)
    return pd.astype(np.uint8).reshape(-1)


@lru_cache(maxsize=256)
def _recipe_subtasks(recipe: str, world: Tuple[Tuple[str, ...], ...], max_num_subtasks: int) -> Tuple[str, ...]:
    """One recipe's subtasks: the set from the planner, in the reference's PYTHONHASHSEED=0 order."""
    try:
        d = recipe_planner.derive_recipe_subtasks(recipe, world, max_num_subtasks)
    except ValueError as ex:
        raise LevelError(str(ex))
    gets = [s for s in d.subtasks if s.startswith("Get(")]
    if gets:      # the plan has to fetch an ingredient the level does not hold; the reference has no goal objects for Get
        raise LevelError("recipe %s needs %s but the level holds no such object" % (recipe, ", ".join(g[4:-1] for g in gets)))
    recorded = levels_data.SUBTASKS.get((recipe,))
    if recorded is None:
        return tuple(d.subtasks)
    try:
        return tuple(recipe_planner.order_like(d.subtasks, recorded))
    except ValueError as ex:
        raise LevelError("recipe %s: %s (regenerate levels_data.py with tools/gen_levels_data.py)" % (recipe, ex))


def derive_level_subtasks(recipes: Sequence[str], obj_contents: Sequence[int], max_num_subtasks: int = 14) -> List[str]:
    """``all_subtasks`` of a level (overcooked_environment.py:452-459): per recipe, the union of the actions on all
    shortest STRIPS plans from the level's objects to the delivered dish, concatenated over the recipe list."""
    world = tuple(tuple(n for n in _NAME_ORDER if CONTENT_BIT[n] & bits) for bits in obj_contents)
    out: List[str] = []
    for r in recipes:
        out += _recipe_subtasks(r, world, max_num_subtasks)
    recorded = levels_data.SUBTASKS.get(tuple(recipes))
    if recorded is not None and list(recorded) != out:
        raise LevelError("derived subtasks of %r differ from the recorded table (regenerate levels_data.py with "
                         "tools/gen_levels_data.py): %r vs %r" % (tuple(recipes), out, list(recorded)))
    return out


def compile_level(level: str, num_agents: int, level_text: Optional[str] = None,
                  subtasks: Optional[Sequence[str]] = None, max_num_subtasks: int = 14) -> CompiledLevel:
    if level_text is None:
        if level not in levels_data.LEVELS:
            raise LevelError("unknown level %r (known: %s)" % (level, ", ".join(sorted(levels_data.LEVELS))))
        level_text = levels_data.LEVELS[level]
    if not 2 <= num_agents <= MAX_AGENTS:
        raise LevelError("num_agents must be 2..%d" % MAX_AGENTS)

    rows: List[str] = []
    recipes: List[str] = []
    starts_xy: List[Tuple[int, int]] = []
    random_reps: List[str] = []
    phase = 1
    for line in level_text.split("\n"):            # lines are only strip('\n')-ed; trailing spaces are tiles (:107)
        if line == "":
            phase += 1
        elif phase == 1:
            rows.append(line)
        elif phase == 2:
            recipes.append(line)
        elif phase == 3:
            if len(starts_xy) < num_agents:
                a, b = line.split(" ")
                starts_xy.append((int(a), int(b)))
        elif phase == 4:
            reps = [ch for ch in line if ch in "tlop"]
            if reps and random_reps:
                # the reference starts a fresh `occupied` set per phase-4 LINE (:157-166), so objects of different lines
                # may share a counter; the device draws all random objects without replacement.  No shipped level has
                # a second line -- refuse instead of silently diverging.
                raise LevelError("more than one line of random (phase-4) objects is not supported: the reference places "
                                 "each line independently, so their objects may collide on one counter")
            random_reps += reps
    if not rows:
        raise LevelError("empty level map")
    height = len(rows)
    width = len(rows[-1])                          # world.width = x + 1 of the last map line (:176)
    if any(len(r) != width for r in rows):
        raise LevelError("ragged level map")
    if width * height > MAX_CELLS:
        raise LevelError("level larger than %d cells" % MAX_CELLS)
    if len(starts_xy) < num_agents:
        raise LevelError("level has %d start positions, %d agents requested" % (len(starts_xy), num_agents))

    tiles = np.zeros(width * height, dtype=np.uint8)
    obj_contents: List[int] = []
    obj_cell: List[int] = []
    for y, row in enumerate(rows):
        for x, ch in enumerate(row):
            c = y * width + x
            if ch in REP_BIT:                      # object on a Counter (:115-122)
                tiles[c] = TILE_COUNTER
                obj_contents.append(REP_BIT[ch])
                obj_cell.append(c)
            else:
                tiles[c] = REP_TILE.get(ch, TILE_FLOOR)
    for ch in random_reps:                         # phase 4 (:157-173)
        obj_contents.append(REP_BIT[ch])
        obj_cell.append(-1)
    if not 1 <= len(obj_contents) <= MAX_OBJECTS:
        raise LevelError("level must hold 1..%d objects" % MAX_OBJECTS)
    foods = [b for b in obj_contents if b != 8]
    if len(set(foods)) != len(foods):
        raise LevelError("each Food may appear at most once in a level (the reference indexes "
                         "list(set(locations))[0], overcooked_environment.py:288,380)")

    for r in recipes:
        if r not in RECIPE_FOODS:
            raise LevelError("unknown recipe %r" % r)
    if subtasks is None:
        subtasks = derive_level_subtasks(recipes, obj_contents, max_num_subtasks)
    subtasks = list(subtasks)
    if not 1 <= len(subtasks) <= MAX_SUBTASKS:
        raise LevelError("the recipe list %r has %d subtasks; the packed state holds one bit per subtask and supports "
                         "1..%d subtasks (every shipped level has at most 9; lists that combine OnionSalad with another "
                         "recipe have 32-44)" % (tuple(recipes), len(subtasks), MAX_SUBTASKS))
    parsed = [parse_subtask(s) for s in subtasks]
    if not any(k == SUBTASK_KIND["Deliver"] for k, _, _ in parsed):
        raise LevelError("no delivery subtask")    # the reference asserts (:251)

    # shaping item list: Plate, then recipes[0].contents sorted by name (:319-321, recipe.py:29)
    r0 = RECIPE_FOODS[recipes[0]]
    items = [8] + [CONTENT_BIT[n] for n in _NAME_ORDER if n != "Plate" and CONTENT_BIT[n] & r0]

    starts = []
    for x, y in starts_xy:
        if not (0 <= x < width and 0 <= y < height) or tiles[y * width + x] != TILE_FLOOR:
            raise LevelError("agent start (%d, %d) is not a floor tile" % (x, y))
        starts.append(y * width + x)

    lv = CompiledLevel(name=level, width=width, height=height, tiles=tiles, starts=starts,
                       object_contents=obj_contents, object_cell=obj_cell, recipes=tuple(recipes),
                       subtasks=subtasks)
    lv.subtask_kind = [k for k, _, _ in parsed]
    lv.subtask_goal = [g for _, g, _ in parsed]
    lv.subtask_arg0 = [a for _, _, a in parsed]
    lv.items = items
    lv.max_path = 2 * (width + height) + 1          # perimeter + 1 (:178, :274)
    lv.path_dist = path_distance_table(tiles, width, height)
    return lv
