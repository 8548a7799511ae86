"""GPU parity, part 2: the CUDA path against the C and Python oracles on fresh inputs at batch
scale -- goal-chasing action streams (so chop / merge / deliver happen), uniform random streams
with device-side auto-reset at BASELINE.json's env counts, the fused Philox rollout, and the
size-independent properties of the path."""
import argparse

import numpy as np
import pytest
import torch

from gym_comm_b200 import levels_data
from gym_comm_b200.vec_env import OvercookedVecEnv
from oracle.c_oracle import COracle
from oracle.drivers import GoalChaser
from oracle.spec_model import SpecEnv

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

ALLERGIC_EGO = dict(CAN_MOVE=False, ALLERGIC=True, BLIND=False)
BLIND_PARTNER = dict(CAN_MOVE=True, ALLERGIC=False, BLIND=True)
DEF = dict(CAN_MOVE=True, ALLERGIC=False, BLIND=False)

CONFIGS = {
    "cfg2": dict(level="open-divider_tomato", num_agents=2, max_num_timesteps=500, communication_on=True,
                 num_communication=10, ego_led=False, fow_radius=2, ego_config=DEF, partner_config=DEF),
    "cfg3": dict(level="partial-divider_salad", num_agents=3, max_num_timesteps=500, communication_on=True,
                 num_communication=10, ego_led=False, fow_radius=2, ego_config=DEF, partner_config=DEF),
    "cfg3_full": dict(level="full-divider_salad", num_agents=3, max_num_timesteps=300, communication_on=True,
                      num_communication=10, ego_led=False, fow_radius=2, ego_config=DEF, partner_config=DEF),
    "cfg4": dict(level="random-open-divider_salad_small_cramped", num_agents=2, max_num_timesteps=900,
                 communication_on=True, num_communication=8, ego_led=False, fow_radius=10,
                 ego_config=ALLERGIC_EGO, partner_config=BLIND_PARTNER),
    "cfg5": dict(level="random-salad-superwide", num_agents=2, max_num_timesteps=900, communication_on=True,
                 num_communication=100, ego_led=False, fow_radius=2, ego_config=DEF, partner_config=DEF),
    "tl3_oddrow": dict(level="open-divider_tl", num_agents=3, max_num_timesteps=200, communication_on=True,
                       num_communication=5, ego_led=True, fow_radius=3, ego_config=DEF, partner_config=DEF),
    # T > 1023: the timestep table no longer fits the shared-memory blob -> computed with an f64 division
    "longT": dict(level="open-divider_tomato", num_agents=2, max_num_timesteps=3001, communication_on=True,
                  num_communication=10, ego_led=False, fow_radius=2, ego_config=DEF, partner_config=DEF),
    # wide rows with 3 / 4 observers: float rows of 16 / 8 envs per pass (multi-pass bulk stores)
    "wide3_c60": dict(level="partial-divider_salad", num_agents=3, max_num_timesteps=40, communication_on=True,
                      num_communication=60, ego_led=False, fow_radius=2, ego_config=DEF, partner_config=DEF),
    "wide4_c100": dict(level="open-divider_salad", num_agents=4, max_num_timesteps=30, communication_on=True,
                       num_communication=100, ego_led=True, fow_radius=1, ego_config=DEF, partner_config=DEF),
    "salad4_commoff": dict(level="open-divider_salad", num_agents=4, max_num_timesteps=150, communication_on=False,
                           num_communication=7, ego_led=False, fow_radius=1, ego_config=DEF, partner_config=DEF),
}


def level_and_subtasks(cfg):
    text = levels_data.LEVELS[cfg["level"]]
    return text, levels_data.SUBTASKS[tuple(text.split("\n\n")[1].split("\n"))]


def oracle_kwargs(cfg):
    return {k: v for k, v in cfg.items() if k != "level"}


def make_gpu(cfg, E, auto_reset, seed=0):
    return OvercookedVecEnv(argparse.Namespace(**cfg), num_envs=E, device=DEV, seed=seed, auto_reset=auto_reset)


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4", "tl3_oddrow", "salad4_commoff"])
def test_goal_chasing_streams_vs_python_oracle(name):
    """E envs, each driven by its own noisy goal chaser; explicit resets with oracle placements."""
    cfg = CONFIGS[name]
    text, subtasks = level_and_subtasks(cfg)
    n, E, T = cfg["num_agents"], 96, 220
    env = make_gpu(cfg, E, auto_reset=False)
    kw = oracle_kwargs(cfg)
    probe = SpecEnv.__new__(SpecEnv)
    probe.n = n
    probe._parse_level(text)
    rng = np.random.default_rng(11)

    def draw():
        if not probe.random_reps:
            return None
        idx = rng.choice(len(probe.counters), len(probe.random_reps), replace=False)
        return [probe.counters[i] for i in idx]

    def cells(p):
        return [x + y * probe.W for x, y in p]

    pls = [draw() for _ in range(E)]
    specs = [SpecEnv(text, subtasks, placements=pls[e], **kw) for e in range(E)]
    chasers = [GoalChaser(specs[e], seed=500 + e, p_random=0.2) for e in range(E)]
    R = len(probe.random_reps)
    pl_t = torch.tensor([cells(p) for p in pls], dtype=torch.int32, device=DEV) if R else None
    obs = env.reset(placements=pl_t)
    events = 0
    for t in range(T):
        acts = np.zeros((E, n, 2), dtype=np.int32)
        for e in range(E):
            navs, comms = chasers[e].act()
            acts[e, :, 0], acts[e, :, 1] = navs, comms
        obs, rew, done = env.step(torch.from_numpy(acts).to(DEV), want_f64=True)
        o, r64, d = obs.cpu().numpy(), env.rewards64.cpu().numpy(), done.cpu().numpy()
        mask = np.zeros(E, dtype=np.uint8)
        newpl = np.zeros((E, max(R, 1)), dtype=np.int32)
        for e in range(E):
            r, dd, sp = specs[e].step(list(acts[e, :, 0]), list(acts[e, :, 1]))
            events += int(sp != 0)
            assert r == r64[e] and dd == bool(d[e]), (name, t, e, r, r64[e])
            want = np.array([specs[e].flat_obs(k) for k in range(n)], dtype=np.float32)
            assert np.array_equal(o[e], want), (name, t, e)
            if dd:
                p = draw()
                specs[e].reset(p)
                chasers[e].on_reset()
                mask[e] = 1
                if p is not None:
                    newpl[e] = cells(p)
        if mask.any():
            obs = env.reset(mask=torch.from_numpy(mask).to(DEV),
                            placements=torch.from_numpy(newpl).to(DEV) if R else None)
            o = obs.cpu().numpy()
            for e in range(E):
                want = np.array([specs[e].flat_obs(k) for k in range(n)], dtype=np.float32)
                assert np.array_equal(o[e], want), (name, "reset", t, e)
    assert events > 20, "chasers should have produced recipe progress"
    env.close()


@pytest.mark.parametrize("name,E,T", [("cfg2", 65536, 1100), ("cfg3", 262144, 120), ("cfg4", 65536, 1000),
                                      ("cfg5", 131072, 60), ("tl3_oddrow", 4099, 450), ("longT", 2000, 3100),
                                      ("wide3_c60", 4099, 90), ("wide4_c100", 3001, 70)])
def test_random_streams_autoreset_vs_c_oracle(name, E, T):
    """BASELINE.json env counts; uniform random actions; auto-reset on the device (random levels
    draw their placements from the shared Philox spec); every step: f64 reward, done, and all
    observations identical to the C oracle; every env crosses at least one reset where T allows."""
    cfg = CONFIGS[name]
    text, subtasks = level_and_subtasks(cfg)
    n = cfg["num_agents"]
    seed = 4242
    env = make_gpu(cfg, E, auto_reset=True, seed=seed)
    ora = COracle(text, subtasks, E, seed=seed, **oracle_kwargs(cfg))
    gen = torch.Generator(device=DEV)
    gen.manual_seed(1)
    term = torch.full((E, n, env.obs_width), -7.0, device=DEV)
    term_o = np.full((E, n, env.obs_width), -7.0)
    check_every = 1 if E <= 70000 else 3
    staggered = T <= cfg["max_num_timesteps"]
    if staggered:
        # the full-size batches run fewer steps than one episode lasts: start env e at clock T_max - 1 - (e mod 50), on
        # the device through oc_get_state / oc_set_state and in the oracle, so that EVERY env crosses an episode
        # boundary (time limit, in-place reset, fresh random placements) within the first 50 steps
        clocks = (cfg["max_num_timesteps"] - 1 - (np.arange(E) % 50)).astype(np.uint32)
        st0 = env.get_state()
        st0[:, 0] = (st0[:, 0] & ~0xFFFF) | torch.from_numpy(clocks.astype(np.int32)).to(DEV)
        env.set_state(st0)
        ora.set_clocks(clocks)
    for t in range(T):
        a = torch.stack([torch.randint(0, 4, (E, n), generator=gen, device=DEV, dtype=torch.int32),
                         torch.randint(0, cfg["num_communication"], (E, n), generator=gen, device=DEV, dtype=torch.int32)], -1).contiguous()
        obs, rew, done = env.step(a, term_obs_out=term, want_f64=True)
        oo, orr, od = ora.step(a.cpu().numpy(), auto_reset=True, term_obs=term_o)
        assert torch.equal(env.rewards64.cpu(), torch.from_numpy(orr)), (name, t)
        assert torch.equal(done.cpu(), torch.from_numpy(od)), (name, t)
        assert torch.equal(rew.cpu()[:, 0], torch.from_numpy(orr.astype(np.float32))), (name, t)
        if t % check_every == 0 or t == T - 1:
            assert torch.equal(obs.cpu(), torch.from_numpy(oo.astype(np.float32))), (name, t)
    assert torch.equal(term.cpu(), torch.from_numpy(term_o.astype(np.float32))), "terminal observations"
    st = env.decode_state()
    os_ = ora.state()
    assert np.array_equal(st["t"], os_["t"])
    assert np.array_equal(st["episodes"], os_["episodes"])
    assert np.array_equal(st["last_completed"], os_["last_completed"])
    if T > cfg["max_num_timesteps"] or (staggered and T >= 51):
        assert st["episodes"].min() >= 1
    env.close()
    ora.close()


@pytest.mark.parametrize("name,E", [("cfg2", 65536), ("cfg4", 20000), ("cfg3_full", 9000), ("cfg5", 10007),
                                    ("wide4_c100", 1501)])
def test_fused_rollout_vs_c_oracle(name, E):
    """oc_rollout (one launch, Philox actions on the device) == the C oracle's twin rollout:
    identical actions, rewards (f32 of the f64), dones and observations, across chunks."""
    cfg = CONFIGS[name]
    text, subtasks = level_and_subtasks(cfg)
    n = cfg["num_agents"]
    env = make_gpu(cfg, E, auto_reset=True, seed=77)
    ora = COracle(text, subtasks, E, seed=77, **oracle_kwargs(cfg))
    F = env.obs_width
    chunk = 12
    obs = torch.zeros((chunk, E, n, F), device=DEV)
    rew = torch.zeros((chunk, E, n), device=DEV)
    done = torch.zeros((chunk, E), dtype=torch.uint8, device=DEV)
    acts = torch.zeros((chunk, E, n, 2), dtype=torch.int32, device=DEV)
    nchunks = 100 if E <= 20000 else 50
    for c in range(nchunks):
        env.rollout(chunk, obs_out=obs, rew_out=rew, done_out=done, actions_out=acts)
        full = (c % 8 == 0) or c == nchunks - 1
        oo, orr, od, oa = ora.rollout(chunk, want_obs=full, want_actions=True)
        assert torch.equal(acts.cpu(), torch.from_numpy(oa)), (name, c)
        assert torch.equal(done.cpu(), torch.from_numpy(od)), (name, c)
        assert torch.equal(rew.cpu()[:, :, 0], torch.from_numpy(orr.astype(np.float32))), (name, c)
        if full:
            assert torch.equal(obs.cpu(), torch.from_numpy(oo.astype(np.float32))), (name, c)
    assert np.array_equal(env.decode_state()["episodes"], ora.state()["episodes"])
    env.close()
    ora.close()


def test_properties_full_size():
    """Size-independent properties at cfg2's full size: determinism, state round trip, one-hot
    structure of the message features, auto-reset == step + masked reset."""
    cfg = CONFIGS["cfg2"]
    E, n = 65536, 2
    a_env = make_gpu(cfg, E, auto_reset=True, seed=3)
    b_env = make_gpu(cfg, E, auto_reset=False, seed=3)
    gen = torch.Generator(device=DEV)
    gen.manual_seed(5)
    lay = a_env.obs_layout
    for t in range(520):
        a = torch.stack([torch.randint(0, 4, (E, n), generator=gen, device=DEV, dtype=torch.int32),
                         torch.randint(0, 10, (E, n), generator=gen, device=DEV, dtype=torch.int32)], -1).contiguous()
        oa, ra, da = a_env.step(a)
        ob, rb, db = b_env.step(a)
        assert torch.equal(ra, rb) and torch.equal(da, db)
        if db.any():
            ob = b_env.reset(mask=db)
        assert torch.equal(oa, ob), t
        if t % 50 == 0:
            assert torch.all(oa[..., lay["agent1_comm"]].sum(-1) == 1) and torch.all(oa[..., lay["agent2_comm"]].sum(-1) == 1)
            assert torch.all(torch.isfinite(ra))
            ts = oa[..., lay["timestep"]]
            assert ts.min() >= 0 and ts.max() <= 1
    # state round trip: export, scramble, import -> identical continuation
    st = a_env.get_state()
    a2 = a_env.step(a)[0].clone()
    a_env.set_state(st)
    assert torch.equal(a_env.get_state(), st)
    assert torch.equal(a_env.step(a)[0], a2)
    a_env.close()
    b_env.close()


@pytest.mark.parametrize("E", [1, 31, 33, 64, 257])
def test_ragged_batch_sizes(E):
    cfg = CONFIGS["cfg5"]
    text, subtasks = level_and_subtasks(cfg)
    env = make_gpu(cfg, E, auto_reset=True, seed=9)
    ora = COracle(text, subtasks, E, seed=9, **oracle_kwargs(cfg))
    rng = np.random.default_rng(E)
    guard = torch.full((E + 1, 2, env.obs_width), 123.0, device=DEV)
    for t in range(40):
        a = np.stack([rng.integers(0, 4, (E, 2)), rng.integers(0, 100, (E, 2))], -1).astype(np.int32)
        obs, rew, done = env.step(torch.from_numpy(a).to(DEV), obs_out=guard[:E], want_f64=True)
        oo, orr, od = ora.step(a, auto_reset=True)
        assert np.array_equal(obs.cpu().numpy(), oo.astype(np.float32))
        assert np.array_equal(env.rewards64.cpu().numpy(), orr)
    assert torch.all(guard[E] == 123.0), "wrote past the last env row"
    env.close()
    ora.close()


def test_out_of_range_and_errors():
    cfg = dict(CONFIGS["cfg2"])
    env = make_gpu(cfg, 8, auto_reset=False)
    with pytest.raises(ValueError):
        env.step(torch.zeros((8, 2, 2), dtype=torch.int64, device=DEV))
    with pytest.raises(ValueError):
        env.step(torch.zeros((7, 2, 2), dtype=torch.int32, device=DEV))
    a = torch.zeros((8, 2, 2), dtype=torch.int32, device=DEV)
    a[:, :, 1] = 999           # out-of-range message -> zero vector (the reference raises IndexError)
    obs, _, _ = env.step(a)
    assert obs[..., env.obs_layout["agent1_comm"]].sum() == 0
    env.close()
    bad = dict(cfg, level="no-such-level")
    with pytest.raises(Exception):
        make_gpu(bad, 4, False)
    bad = dict(cfg, max_num_timesteps=0)
    with pytest.raises(RuntimeError):
        make_gpu(bad, 4, False)


def test_masked_reset_draws_same_placements_as_c_oracle():
    """Random level, auto-reset OFF: explicit masked resets with device-drawn placements (Philox keyed
    by env and episode) must land the objects on the same counters as the C oracle's twin."""
    cfg = dict(CONFIGS["cfg4"], max_num_timesteps=25)
    text, subtasks = level_and_subtasks(cfg)
    E = 3000
    env = make_gpu(cfg, E, auto_reset=False, seed=11)
    ora = COracle(text, subtasks, E, seed=11, **oracle_kwargs(cfg))
    assert np.array_equal(env.reset().cpu().numpy(), ora.reset().astype(np.float32))
    rng = np.random.default_rng(2)
    for t in range(80):
        a = np.stack([rng.integers(0, 4, (E, 2)), rng.integers(0, 8, (E, 2))], -1).astype(np.int32)
        obs, rew, done = env.step(torch.from_numpy(a).to(DEV), want_f64=True)
        oo, orr, od = ora.step(a, auto_reset=False)
        assert np.array_equal(done.cpu().numpy(), od) and np.array_equal(env.rewards64.cpu().numpy(), orr)
        # reset a random subset of the finished envs plus a few unfinished ones
        mask = ((od == 1) | (rng.random(E) < 0.01)).astype(np.uint8)
        if mask.any():
            og = env.reset(mask=torch.from_numpy(mask).to(DEV))
            oc = ora.reset(mask=mask)
            assert np.array_equal(og.cpu().numpy(), oc.astype(np.float32)), t
    st, os_ = env.decode_state(), ora.state()
    assert np.array_equal(st["episodes"], os_["episodes"]) and np.array_equal(st["t"], os_["t"])
    env.close()
    ora.close()


def test_two_handles_with_different_levels_coexist():
    """No global state: handles with different levels / widths / CTA shapes interleave their steps."""
    names = ["cfg5", "cfg2", "cfg3_full"]
    sizes = [4096, 100, 1000]
    envs, oras = [], []
    for nm, E in zip(names, sizes):
        cfg = CONFIGS[nm]
        text, subtasks = level_and_subtasks(cfg)
        envs.append(make_gpu(cfg, E, auto_reset=True, seed=21))
        oras.append(COracle(text, subtasks, E, seed=21, **oracle_kwargs(cfg)))
    rng = np.random.default_rng(8)
    for t in range(60):
        for nm, E, env, ora in zip(names, sizes, envs, oras):
            cfg = CONFIGS[nm]
            n = cfg["num_agents"]
            a = np.stack([rng.integers(0, 4, (E, n)), rng.integers(0, cfg["num_communication"], (E, n))], -1).astype(np.int32)
            obs, rew, done = env.step(torch.from_numpy(a).to(DEV), want_f64=True)
            oo, orr, od = ora.step(a, auto_reset=True)
            assert np.array_equal(obs.cpu().numpy(), oo.astype(np.float32)), (nm, t)
            assert np.array_equal(env.rewards64.cpu().numpy(), orr), (nm, t)
    for env, ora in zip(envs, oras):
        env.close()
        ora.close()


def test_rollout_without_observations():
    """oc_rollout with obs = NULL (state-only fast-forward): rewards/dones/state still match the oracle,
    and a following observed step sees the same world."""
    cfg = CONFIGS["cfg2"]
    text, subtasks = level_and_subtasks(cfg)
    E = 5000
    env = make_gpu(cfg, E, auto_reset=True, seed=31)
    ora = COracle(text, subtasks, E, seed=31, **oracle_kwargs(cfg))
    rew = torch.zeros((700, E, 2), device=DEV)
    done = torch.zeros((700, E), dtype=torch.uint8, device=DEV)
    env.rollout(700, rew_out=rew, done_out=done)
    _, orr, od, _ = ora.rollout(700, want_obs=False)
    assert torch.equal(done.cpu(), torch.from_numpy(od))
    assert torch.equal(rew.cpu()[:, :, 0], torch.from_numpy(orr.astype(np.float32)))
    a = torch.zeros((E, 2, 2), dtype=torch.int32, device=DEV)
    obs, _, _ = env.step(a)
    oo, _, _ = ora.step(a.cpu().numpy(), auto_reset=True)
    assert torch.equal(obs.cpu(), torch.from_numpy(oo.astype(np.float32)))
    env.close()
    ora.close()


def test_replay_of_an_action_sequence_matches_stepping():
    """oc_replay (n steps of caller-given actions in one launch) == n oc_step calls == the C oracle."""
    cfg = CONFIGS["cfg3_full"]
    text, subtasks = level_and_subtasks(cfg)
    E, n, T = 3001, 3, 350
    env_a = make_gpu(cfg, E, auto_reset=True, seed=41)
    env_b = make_gpu(cfg, E, auto_reset=True, seed=41)
    ora = COracle(text, subtasks, E, seed=41, **oracle_kwargs(cfg))
    gen = torch.Generator(device=DEV).manual_seed(41)
    acts = torch.stack([torch.randint(0, 4, (T, E, n), generator=gen, device=DEV, dtype=torch.int32),
                        torch.randint(0, 10, (T, E, n), generator=gen, device=DEV, dtype=torch.int32)], -1).contiguous()
    F = env_a.obs_width
    obs = torch.zeros((T, E, n, F), device=DEV)
    rew = torch.zeros((T, E, n), device=DEV)
    done = torch.zeros((T, E), dtype=torch.uint8, device=DEV)
    env_a.replay(acts, obs_out=obs, rew_out=rew, done_out=done)
    a_cpu = acts.cpu().numpy()
    for t in range(T):
        o_b, r_b, d_b = env_b.step(acts[t])
        assert torch.equal(o_b, obs[t]) and torch.equal(r_b, rew[t]) and torch.equal(d_b, done[t]), t
        oo, orr, od = ora.step(a_cpu[t], auto_reset=True)
        if t % 25 == 0 or t == T - 1:
            assert torch.equal(obs[t].cpu(), torch.from_numpy(oo.astype(np.float32))), t
        assert torch.equal(done[t].cpu(), torch.from_numpy(od)), t
    assert torch.equal(env_a.get_state(), env_b.get_state())
    for e in (env_a, env_b):
        e.close()
    ora.close()
