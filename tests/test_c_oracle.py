"""The C restatement (oracle/oc_oracle.c) must reproduce every golden trace recorded from the
live reference, and agree with the Python restatement on fresh runs (incl. its own RNG resets)."""
import numpy as np
import pytest

from oracle.c_oracle import COracle
from oracle.drivers import GoalChaser
from oracle.spec_model import SpecEnv
from tests.golden_util import env_kwargs, golden_names, load_golden


def _cells(pl, W):
    return np.array([x + y * W for x, y in pl], dtype=np.int32)


@pytest.mark.parametrize("name", golden_names())
def test_c_oracle_replays_golden(name):
    meta, g = load_golden(name)
    n = meta["num_agents"]
    E = 3
    env = COracle(meta["level_text"], meta["subtasks"], E, **env_kwargs(meta))
    W = len(meta["level_text"].split("\n")[0])
    pl = g["placements"]

    def reset(ep):
        p = np.tile(_cells(pl[ep], W), (E, 1)) if pl.shape[1] else None
        return env.reset(placements=p)

    ep = 0
    obs = reset(0)
    for e in range(E):
        assert np.array_equal(obs[e], g["reset_obs"][0])
    for i in range(len(g["navs"])):
        a = np.stack([g["navs"][i].astype(np.int32), g["comms"][i].astype(np.int32)], -1)
        obs, rew, done = env.step(np.tile(a, (E, 1, 1)))
        assert np.all(rew == g["reward"][i]), (i, rew, g["reward"][i])
        assert np.all(done == int(g["done"][i])), i
        st = env.state()
        for e in range(E):
            assert np.array_equal(obs[e], g["obs"][i]), (i, e)
            assert np.array_equal(st["completed"][e], g["completed"][i])
            assert np.array_equal(st["counts"][e], g["counts"][i])
            assert np.array_equal(st["agents"][e], g["agents"][i])
            assert np.array_equal(st["objs"][e], g["objs"][i]), (i, st["objs"][e], g["objs"][i])
        if g["done"][i]:
            ep += 1
            obs = reset(ep)
            for e in range(E):
                assert np.array_equal(obs[e], g["reset_obs"][ep])
    env.close()


@pytest.mark.parametrize("name", ["partial_salad_3a", "cramped_allergic", "custom_level", "open_tl"])
def test_c_oracle_matches_python_oracle_on_fresh_runs(name):
    """Different action streams than the golden traces: a goal chaser per env."""
    meta, _ = load_golden(name)
    kw = env_kwargs(meta)
    n, E, T = meta["num_agents"], 6, 260
    cenv = COracle(meta["level_text"], meta["subtasks"], E, seed=5, **kw)
    probe = SpecEnv.__new__(SpecEnv)
    probe.n = n
    probe._parse_level(meta["level_text"])
    rng = np.random.default_rng(3)

    def draw():
        if not probe.random_reps:
            return None
        idx = rng.choice(len(probe.counters), len(probe.random_reps), replace=False)
        return [probe.counters[i] for i in idx]

    pls = [draw() for _ in range(E)]
    specs = [SpecEnv(meta["level_text"], meta["subtasks"], placements=pls[e], **kw) for e in range(E)]
    if probe.random_reps:
        cenv.reset(placements=np.array([_cells(p, probe.W) for p in pls]))
    chasers = [GoalChaser(specs[e], seed=100 + e, p_random=0.2) for e in range(E)]
    for t in range(T):
        acts = np.zeros((E, n, 2), dtype=np.int32)
        for e in range(E):
            navs, comms = chasers[e].act()
            acts[e, :, 0], acts[e, :, 1] = navs, comms
        obs, rew, done = cenv.step(acts)
        mask = np.zeros(E, dtype=np.uint8)
        newpl = np.zeros((E, max(1, len(probe.random_reps))), dtype=np.int32)
        for e in range(E):
            r, d, _ = specs[e].step(list(acts[e, :, 0]), list(acts[e, :, 1]))
            assert r == rew[e] and d == bool(done[e]), (t, e)
            assert np.array_equal(obs[e], np.array([specs[e].flat_obs(k) for k in range(n)])), (t, e)
            if d:
                p = draw()
                specs[e].reset(p)
                chasers[e].on_reset()
                mask[e] = 1
                if p is not None:
                    newpl[e] = _cells(p, probe.W)
        if mask.any():
            obs = cenv.reset(mask=mask, placements=newpl if probe.random_reps else None)
            for e in range(E):
                assert np.array_equal(obs[e], np.array([specs[e].flat_obs(k) for k in range(n)])), ("reset", t, e)
    cenv.close()
