"""TEST INFRASTRUCTURE ONLY -- CPU restatement (pure Python, single env) of the
reference's environment step + observation path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
may import this.  The product package (``gym_comm_b200``) never does.

Parity status: PINNED.  The reference has no tests or golden vectors of its own
for this path, so the restatement is pinned against
outputs of the reference itself: ``oracle/record_golden.py`` drives the live
reference in the build container and commits traces under ``tests/golden``;
``tests/test_oracle_golden.py`` replays them through this model bit-exactly, and
``tests/test_oracle_vs_reference.py`` lock-steps this model against the live
reference whenever ``/root/reference`` is present.

All ``file:line`` citations are relative to the reference root.

Encoding used throughout the repo (oracle and product agree on it by
documentation, not by shared code):
  content bits   Tomato=1  Lettuce=2  Onion=4  Plate=8   (channel order, core.py:383-388)
  tile codes     0 Floor  1 Counter  2 Cutboard  3 Delivery (core.py:18-26)
  nav actions    0 (0,+1)  1 (0,-1)  2 (-1,0)  3 (+1,0)     (world.py:16)
"""
from __future__ import annotations

from collections import deque
from itertools import combinations

NAV = [(0, 1), (0, -1), (-1, 0), (1, 0)]                     # world.py:16
BIT = {"Tomato": 1, "Lettuce": 2, "Onion": 4, "Plate": 8}    # core.py:383-388 (1 << channel)
REP = {"t": 1, "l": 2, "o": 4, "p": 8}                       # core.py:18-26
FOODS = 7
PLATE = 8
FLOOR, COUNTER, CUTBOARD, DELIVERY = 0, 1, 2, 3
TILE = {"-": COUNTER, "/": CUTBOARD, "*": DELIVERY}
RECIPE_CONTENTS = {                                           # recipe_planner/recipe.py:68-97
    "SimpleTomato": 1, "SimpleLettuce": 2, "Salad": 3, "OnionSalad": 7}
ALPHA = ["Lettuce", "Onion", "Plate", "Tomato"]               # sort order of Object.name parts, core.py:171-175

OBS_KEYS = ["agent1_comm", "agent1_location", "agent2_comm", "agent2_location",
            "agent_is_holding", "completed_subtasks", "is_hidden", "object_encodings_x",
            "object_encodings_y", "state_encodings", "timestep"]


def parse_subtask(s):
    """'Merge(Tomato, Lettuce-Plate)' -> (kind, goal_contents, goal_chopped, arg0_bits).

    Goal template = the object ``get_subtask_obj`` builds (navigation_planner/utils.py:161-209):
    Chop(X) -> ChoppedX; Merge(a, b) -> a U b with every food chopped; Deliver(X) -> X with
    every food chopped.  Equality with a world object is (name, len, full_name)
    (core.py:164-169) == (contents mask, chopped mask)."""
    kind = s[:s.index("(")]
    args = s[s.index("(") + 1:-1].split(", ")
    bits = 0
    for a in args:
        for n in a.split("-"):
            bits |= BIT[n]
    arg0 = 0
    for n in args[0].split("-"):
        arg0 |= BIT[n]
    return kind, bits, bits & FOODS, arg0


class Obj:
    __slots__ = ("contents", "chopped", "loc", "held", "alive", "stamp")

    def __init__(self, contents, loc):
        self.contents = contents
        self.chopped = 0
        self.loc = loc
        self.held = False
        self.alive = True
        self.stamp = 0


class SpecEnv:
    """One Overcooked env with the gym-comm wrapper semantics.

    cfgs = [ego_config, partner_config]; agent 0 gets ego, all others partner
    (overcooked_environment.py:140-143).  ``subtasks`` are the ``str()`` of the
    reference's ``all_subtasks`` in its PYTHONHASHSEED=0 order.
    ``placements`` (random-* levels): list of cells for the phase-4 objects in
    string order (overcooked_environment.py:157-173)."""

    def __init__(self, level_text, subtasks, num_agents=2, max_num_timesteps=500,
                 communication_on=True, num_communication=10, ego_led=False, fow_radius=2,
                 ego_config=None, partner_config=None, placements=None):
        d = {"CAN_MOVE": True, "ALLERGIC": False, "BLIND": False}
        self.cfgs = [dict(d, **(ego_config or {})), dict(d, **(partner_config or {}))]
        self.n = num_agents
        self.T = max_num_timesteps
        self.C = num_communication
        self.comm_on = communication_on
        self.ego_led = ego_led
        self.fow = fow_radius
        self.sub = [parse_subtask(s) for s in subtasks]
        self.S = len(self.sub)
        self._parse_level(level_text)
        # comm index per agent; -1 = all-zero vector.  Initial one-hot at 0 for every
        # agent, even with comm off (overcooked_env.py:89-91); NOT cleared by reset (:284-297).
        self.comm = [0] * max(2, self.n)
        self.reset(placements)

    # ------------------------------------------------------------------ level
    def _parse_level(self, text):
        """overcooked_environment.py:100-178."""
        self.tiles = {}
        self.init_objs = []       # (bits, loc) in world insertion order
        self.random_reps = []     # phase-4 object bits, string order
        self.starts = []
        recipes = []
        phase, y, w = 1, 0, 0
        for line in text.split("\n"):
            if line == "":
                phase += 1
            elif phase == 1:
                for x, ch in enumerate(line):
                    if ch in "tlop":
                        self.tiles[(x, y)] = COUNTER
                        self.init_objs.append((REP[ch], (x, y)))
                    else:
                        self.tiles[(x, y)] = TILE.get(ch, FLOOR)
                w = len(line)
                y += 1
            elif phase == 2:
                recipes.append(line)
            elif phase == 3:
                if len(self.starts) < self.n:
                    a, b = line.split(" ")
                    self.starts.append((int(a), int(b)))
            elif phase == 4:
                self.random_reps += [REP[ch] for ch in line if ch in "tlop"]
        self.W, self.H = w, y
        self.M = 2 * (self.W + self.H) + 1                   # MAX_PATH, :178,:274
        # counters in world.objects["Counter"] order == reading order (:116-126)
        self.counters = [(x, yy) for yy in range(self.H) for x in range(self.W)
                         if self.tiles[(x, yy)] == COUNTER]
        self.delivery = [(x, yy) for yy in range(self.H) for x in range(self.W)
                         if self.tiles[(x, yy)] == DELIVERY]
        # shaping item list: ['Plate'] + recipes[0].contents sorted by name (:319-321, recipe.py:29)
        r0 = RECIPE_CONTENTS[recipes[0]]
        self.items = [PLATE] + [BIT[n] for n in ALPHA if n != "Plate" and (BIT[n] & r0)]
        # floor-to-floor BFS (world.py:61-92 restated; A.6)
        self.floor = {l for l, c in self.tiles.items() if c == FLOOR}
        self.ff = {}
        for src in self.floor:
            dist = {src: 0}
            q = deque([src])
            while q:
                c = q.popleft()
                for a in NAV:
                    nb = (c[0] + a[0], c[1] + a[1])
                    if nb in self.floor and nb not in dist:
                        dist[nb] = dist[c] + 1
                        q.append(nb)
            self.ff[src] = dist

    def pd(self, a, b):
        """World.get_path_distance_between (world.py:114-131)."""
        if a not in self.floor:
            return self.M                       # source node missing -> exception swallowed :126-127
        if b in self.floor:
            return self.ff[a].get(b, self.M)
        best = self.M
        for d in NAV:
            nb = (b[0] + d[0], b[1] + d[1])
            if nb in self.floor and nb in self.ff[a]:
                best = min(best, self.ff[a][nb] + 1)
        return best

    # ------------------------------------------------------------------ reset
    def reset(self, placements=None):
        """OvercookedEnvironment.reset (:180-206); comm buffers untouched."""
        self.t = 0
        self.agents = list(self.starts)
        self.hold = [None] * self.n
        self.objs = []
        self.rank = {}          # name-key (contents mask) -> creation rank in world.objects (world.py:236-237)
        self.next_stamp = 0
        for bits, loc in self.init_objs:
            self._insert(Obj(bits, loc))
        if self.random_reps:
            assert placements is not None and len(placements) == len(self.random_reps)
            for bits, loc in zip(self.random_reps, placements):
                self._insert(Obj(bits, tuple(loc)))
        self.completed = [0] * self.S
        self.count = [0] * self.S

    def _insert(self, o):
        """World.insert: append under key Object.name; key created on first use (world.py:236-237)."""
        self.next_stamp += 1
        o.stamp = self.next_stamp
        if o.contents not in self.rank:
            self.rank[o.contents] = len(self.rank)
        if o not in self.objs:
            self.objs.append(o)

    def ordered(self):
        """Iteration order of world.objects.values() flattened (dict key creation order, then list order)."""
        return sorted((o for o in self.objs if o.alive), key=lambda o: (self.rank[o.contents], o.stamp))

    def _unheld_at(self, loc):
        return [o for o in self.objs if o.alive and not o.held and o.loc == loc]

    # ------------------------------------------------------------------ step
    def step(self, navs, comms):
        """OvercookedMultiEnv.multi_step (overcooked_env.py:207-282) around
        OvercookedEnvironment.step (overcooked_environment.py:211-241).
        Returns (returned_reward f64, done, sparse_reward int)."""
        # comm write (:227-246)
        for k in range(2):
            if not self.comm_on:
                self.comm[k] = -1
            elif k == 1 and self.ego_led:
                self.comm[k] = -1
            else:
                self.comm[k] = int(comms[k])
        # action decode + CAN_MOVE (:248-262)
        acts = [NAV[navs[k]] if self.cfgs[min(k, 1)]["CAN_MOVE"] else (0, 0) for k in range(self.n)]
        self.t += 1                                                          # :213
        # --- check_collisions (:578-613), is_collision (:543-576), ORIGINAL actions for every pair
        nxt = []
        for k in range(self.n):
            cand = (self.agents[k][0] + acts[k][0], self.agents[k][1] + acts[k][1])
            # off-grid asserts in the reference (world.py:314); treated as blocked here (A.1)
            nxt.append(cand if self.tiles.get(cand, COUNTER) == FLOOR else self.agents[k])
        ex = [True] * self.n
        for i, j in combinations(range(self.n), 2):
            if nxt[i] == nxt[j]:
                if nxt[i] == self.agents[i] and acts[i] != (0, 0):
                    ex[j] = False
                elif nxt[j] == self.agents[j] and acts[j] != (0, 0):
                    ex[i] = False
                else:
                    ex[i] = ex[j] = False
            elif self.agents[i] == nxt[j] and self.agents[j] == nxt[i]:
                ex[i] = ex[j] = False
        for k in range(self.n):
            if not ex[k]:
                acts[k] = (0, 0)
        # --- execute_navigation -> interact per agent in order (:615-618, interact.py:4-75)
        for k in range(self.n):
            self._interact(k, acts[k])
        # --- done (:243-270)
        if self.T and self.t >= self.T:
            done = True
        else:
            done = all(self._at_delivery(c, ch) for kind, c, ch, _ in self.sub if kind == "Deliver")
        # --- reward (:399-432)
        rew = 0
        for i, (kind, c, ch, _) in enumerate(self.sub):
            r = 0
            if kind == "Deliver":
                if self._at_delivery(c, ch):
                    r = 3
            else:
                cnt = len({o.loc for o in self.objs if o.alive and o.contents == c and o.chopped == ch})
                if cnt > self.count[i]:
                    r = 1
                self.count[i] = cnt
            rew += r
            if r:
                self.completed[i] = 1
        s0 = self.shaping(0)
        s1 = self.shaping(1)
        return rew - s0 - s1, done, rew                                      # overcooked_env.py:282

    def _interact(self, k, act):
        if act == (0, 0):                                                    # interact.py:12-13
            return
        ax, ay = self.agents[k]
        tgt = (min(max(ax + act[0], 0), self.W - 1), min(max(ay + act[1], 0), self.H - 1))  # world.py:317-320
        tt = self.tiles[tgt]
        h = self.hold[k]
        if tt == FLOOR:                                                      # :19-20, agent.py:311-314
            self.agents[k] = tgt
            if h is not None:
                h.loc = tgt
        elif h is not None:
            here = self._unheld_at(tgt)
            if tt == DELIVERY:                                               # :25-30
                if bin(h.contents).count("1") > 1 and (h.contents & FOODS) == h.chopped:   # core.py:232-237
                    h.loc = tgt
                    h.held = False
                    self.hold[k] = None
            elif here:                                                       # :33-42
                o = here[0]
                if not (h.contents & o.contents & PLATE) \
                        and (h.contents & FOODS) == h.chopped and (o.contents & FOODS) == o.chopped:  # core.py:240-257
                    o.alive = False
                    h.contents |= o.contents
                    h.chopped |= o.chopped
                    self._insert(h)
            else:                                                            # :48-59
                if tt == CUTBOARD and h.contents in (1, 2, 4) and not h.chopped:     # core.py:186-188,201-206
                    h.chopped = h.contents
                else:
                    h.loc = tgt
                    h.held = False
                    self.hold[k] = None
        else:                                                                # :62-75
            here = self._unheld_at(tgt)
            if here and tt != DELIVERY and not self.cfgs[min(k, 1)]["ALLERGIC"]:   # agent.py:296-305
                o = here[0]
                o.held = True
                o.loc = self.agents[k]
                self.hold[k] = o

    def _at_delivery(self, c, ch):
        # only the FIRST Delivery tile is consulted (:259, :402)
        return any(o.alive and o.contents == c and o.chopped == ch and o.loc == self.delivery[0]
                   for o in self.objs)

    # ------------------------------------------------------------------ shaping
    def shaping(self, k):
        """calculate_reward_shaping (:272-397).  Python int/float semantics kept:
        tp stays int 0 until a float term is added; additions in reference order."""
        M = self.M
        a = self.agents[k]
        tp = 0
        U = []
        for i, (kind, c, ch, arg0) in enumerate(self.sub):
            if kind == "Chop" and not self.completed[i]:
                fresh = [o.loc for o in self.objs if o.alive and o.contents == arg0 and o.chopped == 0]
                U.append(self.pd(a, fresh[0]))
        if U:
            tp += ((min(U) + M) + (len(U) - 1) * 2 * M) / M                  # :303-304
        L = {it: [o.loc for o in self.ordered() if o.contents & it] for it in self.items}
        P = []
        for x, y in combinations(self.items, 2):                             # :340-356
            if L[x] and L[y]:
                m = M
                for l1 in L[x]:
                    for l2 in L[y]:
                        m = min(m, self.pd(l1, l2))
                if m != 0:
                    P.append(m)
            else:
                P.append(M)
        if P:                                                                # :359-363
            if tp == 0:
                tp += (min(P) + (len(P) - 1) * M) / M
            else:
                tp += (len(P) * M) / M
        for i, (kind, c, ch, _) in enumerate(self.sub):                      # :370-395
            if kind == "Deliver" and not self.completed[i]:
                D = list({o.loc for o in self.objs if o.alive and o.contents == c and o.chopped == ch})
                if not D:
                    tp += 2
                else:
                    d = self.pd(a, D[0]) + abs(a[0] - D[0][0]) + abs(a[1] - D[0][1])
                    if d == 0:
                        tp += min(self.pd(a, x) + abs(a[0] - x[0]) + abs(a[1] - x[1]) for x in self.delivery) / M
                    else:
                        tp += d / M + 1
        return tp

    # ------------------------------------------------------------------ obs
    def obs(self, k):
        """get_observation2(k, radius=fow_radius) (overcooked_env.py:105-159)."""
        blind = self.cfgs[min(k, 1)]["BLIND"]                                # :115-118
        ax, ay = self.agents[k]
        d = [(0, 0)] * 4
        st = [0] * 4
        hid = [1] * 4
        if not blind:
            for o in self.ordered():                                         # :121-131, last writer wins
                for c in range(4):
                    if o.contents & (1 << c):
                        if c < 3:
                            st[c] = 1 if (o.chopped & (1 << c)) else 0
                        d[c] = (o.loc[0] - ax, o.loc[1] - ay)
            hid = [0 if abs(x) + abs(y) <= self.fow else 1 for x, y in d]     # :133
        vis = [(0, 0) if abs(x) + abs(y) <= self.fow else (x, y) for x, y in d]   # :135

        def cv(i):
            v = [0.0] * self.C
            if self.comm[i] >= 0:
                v[self.comm[i]] = 1.0
            return v
        return {
            "timestep": [self.t / self.T],
            "object_encodings_x": [v[0] for v in vis],
            "object_encodings_y": [v[1] for v in vis],
            "state_encodings": st,
            "is_hidden": hid,
            "completed_subtasks": list(self.completed),
            "agent1_location": [0, 0] if blind else list(self.agents[0]),
            "agent2_location": [0, 0] if blind else list(self.agents[1]),
            # keyed on *ego* BLIND regardless of observer (:154)
            "agent_is_holding": [0, 0] if self.cfgs[0]["BLIND"] else [int(self.hold[k] is not None), 0],
            "agent1_comm": cv(0),
            "agent2_comm": cv(1),
        }

    def flat_obs(self, k):
        o = self.obs(k)
        out = []
        for key in OBS_KEYS:
            out += [float(v) for v in o[key]]
        return out

    # ------------------------------------------------------------------ state
    def full_name(self, o):
        parts = []
        for n in ALPHA:
            b = BIT[n]
            if o.contents & b:
                parts.append(n if n == "Plate" else ("Chopped" if o.chopped & b else "Fresh") + n)
        return "-".join(parts)

    def state_tuple(self):
        """Same canonical form as ``LiveReference.state_tuple``."""
        agents = [(tuple(self.agents[k]), self.full_name(self.hold[k]) if self.hold[k] is not None else None)
                  for k in range(self.n)]
        objs = [(self.full_name(o), tuple(o.loc), bool(o.held)) for o in self.ordered()]
        return (self.t, agents, objs, list(self.completed), list(self.count))
