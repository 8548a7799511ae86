"""Host logic of `OvercookedVecEnv` / `PantheonVecEnv` / `SB3VecEnvAdapter` exercised on CPU through the
test-only emulation backend (tests/emu): argument validation, views, statistics, state decode.
(Parity of the results is the GPU tests' job; this covers the Python plumbing without a GPU.)"""
import argparse

import numpy as np
import pytest
import torch

from gym_comm_b200.pantheon import PantheonVecEnv, SB3VecEnvAdapter
from gym_comm_b200.vec_env import OvercookedVecEnv
from tests.parity_util import EmuVecEnv, emu_library

D = dict(CAN_MOVE=True, ALLERGIC=False, BLIND=False)


def make(E=5, T=6, **kw):
    ns = argparse.Namespace(level="open-divider_tomato", num_agents=2, max_num_timesteps=T, communication_on=True,
                            num_communication=4, ego_led=False, fow_radius=2, ego_config=D, partner_config=D)
    return EmuVecEnv(ns, num_envs=E, device="cpu", lib=emu_library(), **kw)


def test_spaces_layout_and_views():
    env = make()
    assert env.obs_width == 23 + 3 + 2 * 4 and env.action_space.nvec.tolist() == [4, 4]
    assert list(env.observation_space.spaces) == sorted(env.observation_space.spaces)      # gym sorts Dict keys
    obs = env.reset()
    views = env.obs_dict(obs)
    assert sum(v.shape[-1] for v in views.values()) == env.obs_width
    assert views["agent1_comm"].data_ptr() == obs.data_ptr()                                # zero-copy
    assert torch.all(views["agent1_comm"][..., 0] == 1) and torch.all(views["timestep"] == 0)
    assert views["agent1_location"][0, 0].tolist() == [2.0, 1.0] and views["agent2_location"][0, 0].tolist() == [4.0, 1.0]
    env.close()


def test_argument_validation():
    env = make()
    a = torch.zeros((5, 2, 2), dtype=torch.int32)
    with pytest.raises(ValueError):
        env.step(a.to(torch.int64))
    with pytest.raises(ValueError):
        env.step(a[:4])
    with pytest.raises(ValueError):
        env.step(a.transpose(0, 1).contiguous().transpose(0, 1))       # right shape, wrong strides
    with pytest.raises(ValueError):
        env.step(a, obs_out=torch.zeros((5, 2, env.obs_width + 1)))
    with pytest.raises(ValueError):
        env.reset(mask=torch.zeros(5, dtype=torch.bool))
    with pytest.raises(ValueError):
        env.rollout(3, rew_out=torch.zeros((2, 5, 2)))
    env.close()


def test_stats_decode_and_render():
    env = make(E=3, T=4)
    a = torch.zeros((3, 2, 2), dtype=torch.int32)
    a[:, 0, 0] = 3
    for _ in range(4):
        obs, rew, done = env.step(a)
    assert done.tolist() == [1, 1, 1]                       # time limit, auto-reset happened
    st = env.stats()
    assert st["episodes"].tolist() == [1, 1, 1] and st["num_completed_subtasks"].tolist() == [0, 0, 0]
    dec = env.decode_state()
    assert dec["t"].tolist() == [0, 0, 0] and dec["agent_x"][0].tolist() == [2, 4]
    assert (dec["obj_contents"][0] != 0).sum() == 4
    pic = env.render(0).split("\n")
    assert len(pic) == 7 and pic[1].startswith("/   0   1")
    env.close()


def test_pantheon_and_sb3_plumbing():
    class Partner:
        def get_action(self, obs, record=True):
            return torch.zeros((obs.shape[0], 2), dtype=torch.int32)

        def update(self, r, d):
            self.last = (r.clone(), d.clone())

    penv = PantheonVecEnv(make(E=4, T=3), Partner())
    venv = SB3VecEnvAdapter(penv)
    o = venv.reset()
    assert o.shape == (4, penv.obs_dim) and o.dtype == np.float32
    for t in range(3):
        o, r, d, infos = venv.step(np.zeros((4, 2), dtype=np.int64))
    assert d.all() and all("terminal_observation" in i for i in infos)
    assert penv.pop_episode_stats()["episodes"] == 4
    with pytest.raises(ValueError):
        PantheonVecEnv(make(auto_reset=False), Partner())
    with pytest.raises(AttributeError):
        venv.set_attr("x", 1)
    venv.close()


@pytest.mark.parametrize("recurrent", [False, True])
def test_ego_and_partner_learners_run_off_the_env(recurrent):
    """collect_and_train with the feed-forward and the recurrent learner (ego + partner inside the env
    step) on the emulated env: two iterations, finite losses, the partner trained on its own schedule."""
    from gym_comm_b200.pantheon import BatchedOnPolicyAgent, collect_and_train
    from gym_comm_b200.ppo import PPO, PPOConfig, RecurrentPPO
    E = 8
    env = make(E=E, T=5)
    cfg = PPOConfig(n_steps=6, batch_size=24, n_epochs=1)
    mk = (lambda seed: RecurrentPPO(env.obs_width, 4, 4, E, "cpu", cfg, seed=seed, lstm_hidden=8)) if recurrent else \
         (lambda seed: PPO(env.obs_width, 4, 4, E, "cpu", cfg, seed=seed))
    ego, partner = mk(0), BatchedOnPolicyAgent(mk(1))
    penv = PantheonVecEnv(env, partner)
    obs, starts = penv.reset(), torch.ones(E)
    for _ in range(2):
        obs, starts, stats = collect_and_train(penv, ego, obs, starts)
        assert all(np.isfinite(v) for v in stats.values())
    assert partner.iteration >= 1 and ego.n_updates == 2
    assert penv.pop_episode_stats()["episodes"] == E * 2
    env.close()


@pytest.mark.parametrize("flags", [[], ["--recurrent", "--lstm-hidden", "8"]])
def test_train_ppo_main_loop_on_the_emulated_env(flags):
    """train_ppo.main end to end (training iterations, logging, tester.py-style deterministic
    evaluation) with the env swapped for the CPU emulation; learning itself is the GPU test's job."""
    import train_ppo

    def factory(ns, args):
        return EmuVecEnv(ns, num_envs=args.envs, device="cpu", seed=args.seed, auto_reset=True, lib=emu_library())
    hist = train_ppo.main(["--envs", "8", "--n-steps", "5", "--total-timesteps", "81", "--log-every", "1", "--batch-size", "20",
                           "--max-num-timesteps", "4", "--epochs", "1", "--eval-steps", "9", "--device", "cpu"] + flags,
                          env_factory=factory)
    assert len(hist) == 4 and hist[-1]["eval"] and hist[-1]["episodes"] == 8 * 2
    assert all(np.isfinite(v) for h in hist[:3] for v in h["ego_loss"].values())


@pytest.mark.parametrize("E,C", [(5, 4), (67, 7), (33, 100)])
def test_pack_obs_i8_is_the_float_row_without_the_clock(E, C):
    """The compact integer format (oc_pack_obs_i8; the word-per-thread device function run thread by thread
    on the CPU emulation): int8 [E, A, F-1] + f32 [E] rebuild the float rows exactly, for row widths where a
    4-byte output word straddles rows (F-1 = 33, 39, 225) and a batch whose byte count is not a multiple of 4."""
    ns = argparse.Namespace(level="open-divider_tomato", num_agents=2, max_num_timesteps=9, communication_on=True,
                            num_communication=C, ego_led=False, fow_radius=2, ego_config=D, partner_config=D)
    env = EmuVecEnv(ns, num_envs=E, device="cpu", lib=emu_library(), auto_reset=True)
    gen = torch.Generator().manual_seed(E)
    env.reset()
    for t in range(14):
        a = torch.stack([torch.randint(0, 4, (E, 2), generator=gen), torch.randint(0, C, (E, 2), generator=gen)], -1).to(torch.int32)
        obs, _, _ = env.step(a)[:3]
        guard = torch.full((E * 2 * (env.obs_width - 1) + 8,), 99, dtype=torch.int8)
        out = guard[:E * 2 * (env.obs_width - 1)].view(E, 2, env.obs_width - 1)
        i8, ts = env.pack_obs_i8(obs, out=out)
        assert torch.all(guard[-8:] == 99)                                        # nothing written past the end
        assert torch.equal(i8.to(torch.float32), obs[..., :-1])
        assert torch.equal(ts, obs[:, 0, -1]) and torch.equal(ts, obs[:, 1, -1])
    assert i8.min() < 0                                                           # signed deltas survive
    env.close()


def test_terminal_row_gather_on_the_emulation():
    """The device function behind the host path's terminal observations (oc_gather_term_kernel): rows of the
    finished envs land densely, in index order, as float rows or in the compact format."""
    import ctypes as C
    E, Cn = 37, 5
    ns = argparse.Namespace(level="open-divider_tomato", num_agents=2, max_num_timesteps=9, communication_on=True,
                            num_communication=Cn, ego_led=False, fow_radius=2, ego_config=D, partner_config=D)
    env = EmuVecEnv(ns, num_envs=E, device="cpu", lib=emu_library(), auto_reset=True)
    F = env.obs_width
    rng = np.random.default_rng(0)
    term = np.ascontiguousarray(rng.integers(-9, 10, (E, 2, F)).astype(np.float32))
    term[..., -1] = rng.random((E, 1)).astype(np.float32)
    idx = np.array([0, 3, 4, 17, 36], np.int32)
    f = emu_library().lib.emu_gather_term
    f.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3
    out = np.full((len(idx) + 1, 2, F), 7, np.float32)
    f(env._handle, term.ctypes.data, idx.ctypes.data, len(idx), out.ctypes.data, None, None)
    assert np.array_equal(out[:-1], term[idx]) and np.all(out[-1] == 7)
    o8, ts = np.full((len(idx) + 1, 2, F - 1), 7, np.int8), np.zeros(len(idx), np.float32)
    f(env._handle, term.ctypes.data, idx.ctypes.data, len(idx), None, o8.ctypes.data, ts.ctypes.data)
    assert np.array_equal(o8[:-1], term[idx][..., :-1].astype(np.int8)) and np.all(o8[-1] == 7)
    assert np.array_equal(ts, term[idx][:, 0, -1])
    env.close()


@pytest.mark.parametrize("flags", [[], ["--recurrent", "--lstm-hidden", "8"]])
def test_saved_policies_play_through_evaluate_policy(tmp_path, flags):
    """train_ppo --save-dir -> evaluate_policy --ego-load / --alt-load (trainer.py:129-133 -> tester.py:64-128) on the
    emulated env: the loaded learners are the saved ones, the requested number of games is played, and the
    statistics line is consistent."""
    import evaluate_policy
    import train_ppo
    from gym_comm_b200.ppo import load_learner

    def factory(ns, args):
        return EmuVecEnv(ns, num_envs=args.envs, device="cpu", seed=args.seed, auto_reset=True, lib=emu_library())
    d = str(tmp_path / "model")
    train_ppo.main(["--envs", "8", "--n-steps", "5", "--iters", "2", "--log-every", "1", "--batch-size", "20",
                    "--max-num-timesteps", "6", "--epochs", "1", "--device", "cpu", "--save-dir", d] + flags, env_factory=factory)
    a, b = load_learner(d + "/ppo_ego.pt", 4, "cpu"), load_learner(d + "/ppo_ego.pt", 4, "cpu")
    assert type(a).__name__ == ("RecurrentPPO" if flags else "PPO")
    for x, y in zip(a.policy.state_dict().values(), b.policy.state_dict().values()):
        assert torch.equal(x, y)
    out = evaluate_policy.main(["--max-num-timesteps", "6", "--ego-load", d + "/ppo_ego.pt", "--alt-load", d + "/ppo_partner1.pt",
                            "-t", "22", "--envs", "4", "-d", "cpu"], env_factory=factory)
    # exactly the requested games: a quota of 22 // 4 per env, the first 22 % 4 envs one more -- every env plays its
    # share to the end, so long (timed-out) games are not crowded out by short ones that finish and replay
    assert out["episodes"] == 22 and out["env_steps"] % 4 == 0 and out["env_steps"] // 4 >= 6 * 5
    assert out["ep_len_mean"] <= 6 and 0.0 <= out["delivered_frac"] <= 1.0
    assert np.isfinite(out["average_reward"]) and out["standard_deviation"] >= 0.0
    with pytest.raises(ValueError):                                         # a policy of another message width is refused
        evaluate_policy.main(["--max-num-timesteps", "6", "--num-communication", "3", "--ego-load", d + "/ppo_ego.pt",
                          "--alt-load", d + "/ppo_partner1.pt", "-t", "4", "--envs", "4", "-d", "cpu"], env_factory=factory)


def test_episode_statistics_match_a_manual_tally():
    """pop_episode_stats (count, mean / std of the episode return, mean length) against per-episode sums kept by
    hand, with a static random partner (`BatchedStaticPolicyAgent`, never learning)."""
    from gym_comm_b200.pantheon import BatchedStaticPolicyAgent
    from gym_comm_b200.ppo import PPO, PPOConfig
    E = 6
    env = make(E=E, T=7, auto_reset=True)
    partner = BatchedStaticPolicyAgent(PPO(env.obs_width, 4, 4, E, "cpu", PPOConfig(n_steps=4, batch_size=8), seed=5))
    penv = PantheonVecEnv(env, partner)
    gen = torch.Generator().manual_seed(1)
    penv.reset()
    run, rets, lens = np.zeros(E), [], []
    for t in range(30):
        a = torch.stack([torch.randint(0, 4, (E,), generator=gen), torch.randint(0, 4, (E,), generator=gen)], -1).to(torch.int32)
        _, r, d = penv.step(a)
        run += r.numpy()
        for e in np.flatnonzero(d.numpy()):
            rets.append(run[e]); lens.append(7); run[e] = 0.0
    st = penv.pop_episode_stats()
    assert st["episodes"] == len(rets) == E * 4
    assert abs(st["ep_rew_mean"] - np.mean(rets)) < 1e-4 and abs(st["ep_rew_std"] - np.std(rets)) < 1e-3
    last = env.stats()["num_completed_subtasks"].double().mean().item()       # every env has finished episodes by now
    assert st["num_completed_subtasks"] == pytest.approx(last) and 0.0 <= last <= 3.0
    assert st["ep_len_mean"] == 7 and penv.pop_episode_stats()["episodes"] == 0
    assert partner.model.n_updates == 0                                      # a static agent never trains
    env.close()


@pytest.mark.parametrize("level,A", [("open-divider_tomato", 2), ("random-open-divider_salad_small_cramped", 2), ("partial-divider_salad", 3)])
def test_state_injection_continues_identically(level, A):
    """Mid-game states exported from one env and injected into a fresh handle (oc_get_state / oc_set_state): both
    continue with identical observations, rewards, dones and episode counters under the same actions, across
    auto-resets (random placements included: they are keyed by env index and episode number, both in the state)."""
    E, C, T = 45, 4, 13
    ns = argparse.Namespace(level=level, num_agents=A, max_num_timesteps=T, communication_on=True, num_communication=C,
                            ego_led=False, fow_radius=2, ego_config=D, partner_config=D)
    a_env = EmuVecEnv(ns, num_envs=E, device="cpu", seed=9, auto_reset=True, lib=emu_library())
    gen = torch.Generator().manual_seed(2)

    def act():
        return torch.stack([torch.randint(0, 4, (E, A), generator=gen), torch.randint(0, C, (E, A), generator=gen)], -1).to(torch.int32)
    a_env.reset()
    for _ in range(T + 5):                                  # past the first auto-reset
        a_env.step(act())
    st = a_env.get_state()
    b_env = EmuVecEnv(ns, num_envs=E, device="cpu", seed=9, auto_reset=True, lib=emu_library())
    b_env.set_state(st)
    assert torch.equal(b_env.get_state(), st)
    for t in range(2 * T):
        a = act()
        oa, ra, da = (x.clone() for x in a_env.step(a, want_f64=True))
        r64 = a_env.rewards64.clone()
        ob, rb, db = b_env.step(a, want_f64=True)
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db) and torch.equal(r64, b_env.rewards64), t
    sa, sb = a_env.stats(), b_env.stats()
    assert torch.equal(sa["episodes"], sb["episodes"]) and torch.equal(sa["num_completed_subtasks"], sb["num_completed_subtasks"])
    assert torch.equal(a_env.get_state(), b_env.get_state())
    a_env.close()
    b_env.close()
