"""The Python restatement (oracle/spec_model.py) must reproduce every golden trace that
oracle/record_golden.py recorded from the live reference: returned reward (exact f64),
done, every observer's flat observation, and the canonical state, at every step and after
every reset."""
import numpy as np
import pytest

from oracle.spec_model import SpecEnv
from tests.golden_util import env_kwargs, golden_names, load_golden

# rewards of the scripted open-divider_tomato solve, copied from the survey-time
# probe of the live reference (an answer key independent of the recorder).
A9_REWARDS = [
    -8.172413793103448, -8.10344827586207, -8.137931034482758, -8.10344827586207, -8.068965517241379,
    -8.03448275862069, -8.068965517241379, -5.0, -6.0, -6.0, -6.0, -6.0, -6.0, -6.0, -6.0,
    -0.9655172413793105, -1.8275862068965516, -1.6896551724137931, -1.5517241379310347,
    -1.4137931034482758, -1.2758620689655173, -1.2758620689655171, 1.0]


def test_golden_present():
    assert len(golden_names()) >= 12


def test_a9_known_answers():
    meta, g = load_golden("tomato_a9_script")
    assert list(g["reward"]) == A9_REWARDS
    assert list(g["done"]) == [False] * 22 + [True]
    assert list(g["completed"][7]) == [0, 0, 1]     # step 8: chop
    assert list(g["completed"][15]) == [0, 1, 1]    # step 16: merge with plate
    assert list(g["completed"][22]) == [1, 1, 1]    # step 23: deliver


@pytest.mark.parametrize("name", golden_names())
def test_spec_model_replays_golden(name):
    meta, g = load_golden(name)
    n = meta["num_agents"]
    pl = g["placements"]
    env = SpecEnv(meta["level_text"], meta["subtasks"], placements=[tuple(p) for p in pl[0]] or None,
                  **env_kwargs(meta))
    ep = 0
    assert np.array_equal(np.array([env.flat_obs(k) for k in range(n)]), g["reset_obs"][0])
    for i in range(len(g["navs"])):
        r, d, _ = env.step(list(g["navs"][i]), list(g["comms"][i]))
        assert r == g["reward"][i], (i, r, g["reward"][i])
        assert d == bool(g["done"][i]), i
        assert np.array_equal(np.array([env.flat_obs(k) for k in range(n)]), g["obs"][i]), i
        assert list(env.completed) == list(g["completed"][i])
        assert list(env.count) == list(g["counts"][i])
        assert env.t == g["t"][i]
        assert [list(a) for a in env.agents] == g["agents"][i].tolist()
        objs = [[o.contents, o.chopped, o.loc[0], o.loc[1], int(o.held)] for o in env.ordered()]
        assert objs == [row for row in g["objs"][i].tolist() if row[0] >= 0], i
        if d:
            ep += 1
            env.reset([tuple(p) for p in pl[ep]] or None)
            assert np.array_equal(np.array([env.flat_obs(k) for k in range(n)]), g["reset_obs"][ep])
