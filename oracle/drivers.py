"""TEST INFRASTRUCTURE ONLY -- action sources used to exercise the env path.

Random actions almost never reach merge/deliver (0-4 reward
events in 1200 random steps), so parity runs use a *noisy goal chaser*: every
agent repeatedly picks an interesting non-floor tile (one carrying an object, a
cutboard, the delivery tile, or a random counter), walks a BFS shortest path to
an adjacent floor cell and then presses into the tile; a fraction of the
actions is uniformly random.  The chaser reads a ``SpecEnv`` (which runs in
lock-step with whatever is being checked), never the product.
"""
from __future__ import annotations

import random
from collections import deque

from .spec_model import NAV, FLOOR, COUNTER, CUTBOARD, DELIVERY


class GoalChaser:
    def __init__(self, env, seed=0, p_random=0.15):
        self.env = env
        self.rng = random.Random(seed)
        self.p_random = p_random
        self.goal = [None] * env.n
        self.patience = [0] * env.n

    def _pick_goal(self, k):
        e = self.env
        obj_tiles = [o.loc for o in e.objs if o.alive and not o.held and e.tiles[o.loc] != FLOOR]
        special = [l for l, c in e.tiles.items() if c in (CUTBOARD, DELIVERY)]
        counters = [l for l, c in e.tiles.items() if c == COUNTER]
        r = self.rng.random()
        if self.rng.random() < 0.65:
            # purposeful choice: make recipe progress so merges / deliveries get exercised
            h = e.hold[k]
            on_tiles = [o for o in e.objs if o.alive and not o.held and e.tiles[o.loc] != DELIVERY]
            if h is None:
                pool = [o.loc for o in on_tiles]
            elif h.contents in (1, 2, 4) and not h.chopped:
                pool = [l for l, c in e.tiles.items() if c == CUTBOARD and not e._unheld_at(l)]
            elif bin(h.contents).count("1") > 1 and self.rng.random() < 0.6:
                pool = [l for l, c in e.tiles.items() if c == DELIVERY]
            else:
                pool = [o.loc for o in on_tiles if (o.contents & 7) == o.chopped and not (o.contents & h.contents)]
            if pool:
                return self.rng.choice(pool)
        if e.hold[k] is not None:
            pool = special if r < 0.5 else (obj_tiles if r < 0.8 and obj_tiles else counters)
        else:
            pool = obj_tiles if r < 0.7 and obj_tiles else (special if r < 0.8 else counters)
        return self.rng.choice(pool)

    def _nav_towards(self, k, goal):
        e = self.env
        src = e.agents[k]
        # BFS over floor from src; stop at a floor cell adjacent to goal
        prev = {src: None}
        q = deque([src])
        hit = None
        while q:
            c = q.popleft()
            if abs(c[0] - goal[0]) + abs(c[1] - goal[1]) == 1:
                hit = c
                break
            for i, a in enumerate(NAV):
                nb = (c[0] + a[0], c[1] + a[1])
                if e.tiles.get(nb, COUNTER) == FLOOR and nb not in prev:
                    prev[nb] = (c, i)
                    q.append(nb)
        if hit is None:
            return None
        if hit == src:
            d = (goal[0] - src[0], goal[1] - src[1])
            return NAV.index(d)
        c = hit
        while prev[c][0] != src:
            c = prev[c][0]
        return prev[c][1]

    def act(self):
        e = self.env
        navs, comms = [], []
        for k in range(e.n):
            if self.goal[k] is None or self.patience[k] <= 0:
                self.goal[k] = self._pick_goal(k)
                self.patience[k] = self.rng.randint(3, 14)
            self.patience[k] -= 1
            nav = None
            if self.rng.random() >= self.p_random:
                nav = self._nav_towards(k, self.goal[k])
                if nav is not None:
                    a = NAV[nav]
                    if (e.agents[k][0] + a[0], e.agents[k][1] + a[1]) == self.goal[k]:
                        self.goal[k] = None      # pressed into it; choose a new goal next time
            if nav is None:
                nav = self.rng.randrange(4)
            navs.append(nav)
            comms.append(self.rng.randrange(e.C))
        return navs, comms

    def on_reset(self):
        self.goal = [None] * self.env.n


def random_placements(env, rng):
    """Uniform over ALL Counter tiles without replacement, in phase-4 string order
    (overcooked_environment.py:157-173)."""
    return rng.sample(env.counters, len(env.random_reps)) if env.random_reps else None
