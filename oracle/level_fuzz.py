"""TEST INFRASTRUCTURE ONLY -- random kitchen generator for fuzzing the restatements and the device
code against the live reference on layouts the shipped levels never exercise (narrow corridors,
several Delivery tiles / cutboards, unreachable pockets, objects next to each other...)."""
from __future__ import annotations

import random


def random_level(seed: int, num_agents: int = 2):
    rng = random.Random(seed)
    W, H = rng.randint(5, 9), rng.randint(4, 7)
    g = [["-" if x in (0, W - 1) or y in (0, H - 1) else " " for x in range(W)] for y in range(H)]
    interior = [(x, y) for y in range(1, H - 1) for x in range(1, W - 1)]
    for (x, y) in rng.sample(interior, k=rng.randint(0, max(0, len(interior) // 4))):
        g[y][x] = "-"
    floor = [(x, y) for (x, y) in interior if g[y][x] == " "]
    if len(floor) < num_agents + 1:
        return random_level(seed + 7919, num_agents)
    # tiles that touch floor: candidates for stations and objects
    def touches_floor(x, y):
        return any(0 <= x + dx < W and 0 <= y + dy < H and g[y + dy][x + dx] == " "
                   for dx, dy in ((0, 1), (0, -1), (1, 0), (-1, 0)))
    cand = [(x, y) for y in range(H) for x in range(W) if g[y][x] == "-" and touches_floor(x, y)]
    rng.shuffle(cand)
    recipe = rng.choice([["SimpleTomato"], ["Salad"], ["SimpleTomato", "SimpleLettuce"], ["SimpleLettuce"]])
    need = {"SimpleTomato": "t", "SimpleLettuce": "l", "Salad": "tl"}
    foods = sorted(set("".join(need[r] for r in recipe)) | (set("l") if rng.random() < 0.3 else set()) |
                   (set("t") if rng.random() < 0.3 else set()))
    nplates = rng.randint(max(1, len(recipe)), 3)
    stations = ["*"] * rng.randint(1, 2) + ["/"] * rng.randint(1, 2)
    objs = list(foods) + ["p"] * nplates
    if len(objs) > 6:
        objs = objs[:6]
    if len(cand) < len(stations) + len(objs):
        return random_level(seed + 104729, num_agents)
    for ch in stations:
        x, y = cand.pop()
        g[y][x] = ch
    random_objs = rng.random() < 0.35
    if not random_objs:
        for ch in objs:
            x, y = cand.pop()
            g[y][x] = ch
    starts = rng.sample(floor, num_agents)
    text = "\n".join("".join(r) for r in g) + "\n\n" + "\n".join(recipe) + "\n\n" + \
        "\n".join("%d %d" % s for s in starts) + "\n"
    if random_objs:
        rng.shuffle(objs)
        text += "\n" + "".join(objs) + "\n"
    return text
