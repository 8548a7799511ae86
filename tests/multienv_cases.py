"""Cases for the reference-shaped single-env surface (`OvercookedMultiEnv.multi_step / multi_reset /
get_observation2`, gym_comm/envs/overcooked_env.py:105-297) against golden traces of the live
reference: 11-key dict observations, Python-float reward with the reference's f64 value, done.
`envkw` selects the backend: {} = the CUDA library (tests/test_gpu_multienv_api.py), or the CPU emulation of
the device code (tests/test_multienv_api_cpu.py)."""
import numpy as np

from gym_comm_b200.vec_env import OvercookedMultiEnv


def _make(ns, env_cls=None, **kw):
    """`env_cls`: the CUDA-backed product class by default; the CPU suite passes tests.parity_util.EmuMultiEnv."""
    return (env_cls or OvercookedMultiEnv)(ns, **kw)
from tests.golden_util import load_golden
from tests.parity_util import namespace_from_meta

KEYS = ["agent1_comm", "agent1_location", "agent2_comm", "agent2_location", "agent_is_holding",
        "completed_subtasks", "is_hidden", "object_encodings_x", "object_encodings_y", "state_encodings", "timestep"]


def _split_golden(meta, flat):
    C, S = meta["num_communication"], len(meta["subtasks"])
    sizes = [C, 2, C, 2, 2, S, 4, 4, 4, 4, 1]
    out, o = {}, 0
    for k, n in zip(KEYS, sizes):
        out[k] = flat[o:o + n]
        o += n
    return out


TRACES = ["tomato_a9_script", "cramped_allergic", "open_tl"]


def run_multi_step_matches_reference_trace(name, **envkw):
    meta, g = load_golden(name)
    env = _make(namespace_from_meta(meta), level_text=meta["level_text"], subtasks=meta["subtasks"], **envkw)
    W = len(meta["level_text"].split("\n")[0])

    def placements(ep):
        pl = g["placements"][ep]
        return None if pl.shape[0] == 0 else [int(x + y * W) for x, y in pl]

    def check(obs_pair, want_rows):
        for k in range(2):
            want = _split_golden(meta, want_rows[k])
            assert set(obs_pair[k].keys()) == set(KEYS)
            for key in KEYS:
                got = np.asarray(obs_pair[k][key], dtype=np.float64)
                assert np.array_equal(got, want[key]), (name, key, got, want[key])      # timestep too: exact f64

    ep = 0
    check(env.multi_reset(placements(0)), g["reset_obs"][0])
    assert env.action_space.nvec.tolist() == [4, meta["num_communication"]]
    for i in range(min(len(g["navs"]), 300)):
        a0 = (int(g["navs"][i][0]), int(g["comms"][i][0]))
        a1 = (int(g["navs"][i][1]), int(g["comms"][i][1]))
        (o0, o1), (r0, r1), done, info = env.multi_step(a0, a1)
        assert isinstance(r0, float) and r0 == r1 == g["reward"][i], (i, r0, g["reward"][i])
        assert done is bool(g["done"][i]) and info == {}
        check((o0, o1), g["obs"][i])
        if done:
            ep += 1
            check(env.multi_reset(placements(ep)), g["reset_obs"][ep])
    env.close()


def run_known_answer_survey_a7(**envkw):
    """SURVEY A.7 known answer: fresh open-divider_tomato, one multi_step, observer 0, radius 2."""
    import argparse
    d = dict(CAN_MOVE=True, ALLERGIC=False, BLIND=False)
    ns = argparse.Namespace(level="open-divider_tomato", num_agents=2, max_num_timesteps=500, communication_on=True,
                            num_communication=5, ego_led=False, fow_radius=2, ego_config=d, partner_config=d)
    env = _make(ns, **envkw)
    (o0, _), (r, _), done, _ = env.multi_step((3, 2), (0, 4))
    assert r == -8.206896551724139 and done is False
    assert o0["object_encodings_x"].tolist() == [2, 3, 0, 2] and o0["object_encodings_y"].tolist() == [-1, 0, 0, 5]
    assert o0["is_hidden"].tolist() == [1, 1, 0, 1] and o0["state_encodings"].tolist() == [0, 0, 0, 0]
    assert o0["agent1_location"].tolist() == [3, 1] and o0["agent2_location"].tolist() == [4, 2]
    assert o0["agent1_comm"].tolist() == [0, 0, 1, 0, 0] and o0["agent2_comm"].tolist() == [0, 0, 0, 0, 1]
    assert o0["completed_subtasks"].tolist() == [0, 0, 0] and o0["agent_is_holding"].tolist() == [False, False]
    assert float(o0["timestep"][0]) == 0.002
    env.close()


def run_multiagentenv_step_reset_with_partner(**envkw):
    """`step(action)` / `reset()` with an embedded partner, as trainer.py drives the env
    (pantheonrl/common/multiagentenv.py:172-243)."""
    import argparse

    class Partner:
        def __init__(self):
            self.seen, self.updates = [], []

        def get_action(self, obs):
            self.seen.append(obs.obs)
            return (0, 3)

        def update(self, reward, done):
            self.updates.append((reward, done))

    d = dict(CAN_MOVE=True, ALLERGIC=False, BLIND=False)
    ns = argparse.Namespace(level="open-divider_tomato", num_agents=2, max_num_timesteps=5, communication_on=True,
                            num_communication=5, ego_led=False, fow_radius=2, ego_config=d, partner_config=d)
    env = _make(ns, **envkw)
    twin = _make(ns, **envkw)
    p = Partner()
    env.add_partner_agent(p)
    ego_obs = env.reset()
    o_twin = twin.multi_reset()
    assert all(np.array_equal(ego_obs[k], o_twin[0][k]) for k in KEYS)
    prev = ego_obs
    for t in range(5):
        obs, r, done, info = env.step((3, 1))
        (t0, t1), (tr, _), tdone, _ = twin.multi_step((3, 1), (0, 3))
        assert r == tr and done is tdone and info["_partnerid"] == [0]
        if not done:
            assert all(np.array_equal(obs[k], t0[k]) for k in KEYS)
            prev = obs
        else:
            assert all(np.array_equal(obs[k], prev[k]) for k in KEYS)     # previous ego obs on done
    assert done and len(p.seen) == 5
    assert p.updates[0] == (0.0, False) and len(p.updates) == 6 and p.updates[-1][1] is True
    env.close()
    twin.close()


def run_partner_selection_and_n_step(**envkw):
    """set_partnerid / resample policies (multiagentenv.py:103-147), n_step / n_reset (:395-409), cost_fn, render."""
    import argparse
    import contextlib
    import io

    class Const:
        def __init__(self, a):
            self.a, self.calls = a, 0

        def get_action(self, obs):
            self.calls += 1
            return self.a

        def update(self, reward, done):
            pass

    d = dict(CAN_MOVE=True, ALLERGIC=False, BLIND=False)
    ns = argparse.Namespace(level="open-divider_tomato", num_agents=2, max_num_timesteps=3, communication_on=True,
                            num_communication=5, ego_led=False, fow_radius=2, ego_config=d, partner_config=d)
    env = _make(ns, **envkw)
    a, b, c = Const((0, 1)), Const((1, 2)), Const((2, 3))
    for p in (a, b, c):
        env.add_partner_agent(p)
    env.reset()                                            # round robin: 0 -> 1
    _, _, _, info = env.step((3, 0))
    assert info["_partnerid"] == [1] and (a.calls, b.calls, c.calls) == (0, 1, 0)
    env.set_partnerid(2)
    assert env.step((3, 0))[3]["_partnerid"] == [2] and c.calls == 1
    env.reset()                                            # 2 -> 0
    assert env.step((3, 0))[3]["_partnerid"] == [0]
    env.set_resample_policy("random")
    np.random.seed(0)
    seen = set()
    for _ in range(12):
        env.reset()
        seen.add(env.partnerid)
    assert seen == {0, 1, 2}
    env.set_resample_policy("default")
    try:
        env.set_resample_policy("nope")
        raise AssertionError("invalid policy accepted")
    except ValueError:
        pass
    who, obs = env.n_reset()
    assert who == (0, 1) and float(obs[0].obs["timestep"][0]) == 0.0 and obs[1].action_mask is None
    who, obs, rews, done, info = env.n_step([(3, 0), (0, 1)])
    assert who == (0, 1) and rews[0] == rews[1] and done is False and info == {}
    assert obs[0].obs["agent1_location"].tolist() == [3, 1] and obs[1].obs["agent2_comm"].tolist() == [0, 1, 0, 0, 0]
    assert env.cost_fn() == 1
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        env.render()
    assert "agent1_location" in out.getvalue() and len(out.getvalue().splitlines()) >= 7
    env.close()
