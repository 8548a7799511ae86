"""GPU: the batched PantheonRL layer has the reference's step semantics, and a short PPO run off the
GPU env executes end to end with finite losses and improving shaped return."""
import argparse

import numpy as np
import pytest
import torch

from gym_comm_b200.pantheon import BatchedOnPolicyAgent, PantheonVecEnv
from gym_comm_b200.vec_env import OvercookedVecEnv

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ns(T=40):
    d = dict(CAN_MOVE=True, ALLERGIC=False, BLIND=False)
    return argparse.Namespace(level="open-divider_tomato", num_agents=2, max_num_timesteps=T, communication_on=True,
                              num_communication=10, ego_led=False, fow_radius=2, ego_config=d, partner_config=d)


class ScriptedPartner:
    """Partner with a fixed action table; records what PantheonVecEnv feeds it."""

    def __init__(self, table):
        self.table, self.t, self.obs_seen, self.updates = table, 0, [], []

    def get_action(self, obs, record=True):
        self.obs_seen.append(obs.clone())
        a = self.table[self.t]
        self.t += 1
        return a

    def update(self, reward, done):
        self.updates.append((reward.clone(), done.clone()))


def test_pantheon_layer_semantics():
    E, T, steps = 64, 40, 100
    gen = torch.Generator(device=DEV).manual_seed(0)
    acts = torch.stack([torch.randint(0, 4, (steps, E, 2), generator=gen, device=DEV, dtype=torch.int32),
                        torch.randint(0, 10, (steps, E, 2), generator=gen, device=DEV, dtype=torch.int32)], -1)
    partner = ScriptedPartner(acts[:, :, 1])
    penv = PantheonVecEnv(OvercookedVecEnv(_ns(T), num_envs=E, device=DEV, seed=1), partner)
    ref = OvercookedVecEnv(_ns(T), num_envs=E, device=DEV, seed=1)
    o_ref = ref.reset().clone()
    o = penv.reset()
    assert torch.equal(o, o_ref[:, 0])
    for t in range(steps):
        prev_ego = o_ref[:, 0].clone()
        assert torch.equal(partner_obs_expected := o_ref[:, 1], penv._obs[:, 1])
        o, r, d = penv.step(acts[t, :, 0])
        o_ref, r_ref, d_ref = ref.step(acts[t].contiguous())
        o_ref = o_ref.clone()
        assert torch.equal(partner.obs_seen[t], partner_obs_expected)       # partner saw the pre-step obs
        assert torch.equal(o, o_ref[:, 0]) and torch.equal(r, r_ref[:, 0]) and torch.equal(d, d_ref)
        assert torch.equal(partner.updates[t][0], r_ref[:, 1]) and torch.equal(partner.updates[t][1], d_ref)
        if d.any():
            m = d.bool()
            assert torch.equal(penv.terminal_obs[m], prev_ego[m])           # previous ego obs on done
    st = penv.pop_episode_stats()
    assert st["episodes"] == E * (steps // T) and st["ep_len_mean"] == T


def test_short_training_run_improves_return():
    import train_ppo
    hist = train_ppo.main(["--envs", "2048", "--n-steps", "64", "--iters", "60", "--log-every", "10",
                           "--batch-size", "16384", "--max-num-timesteps", "100", "--device", DEV])
    assert len(hist) == 6
    for h in hist:
        assert all(np.isfinite(v) for v in h["ego_loss"].values())
        assert h["episodes"] > 0
    assert hist[-1]["partner_updates"] >= 50
    # dense shaping makes the return move quickly once the agents walk towards the tomato
    assert hist[-1]["ep_rew_mean"] > hist[0]["ep_rew_mean"], (hist[0]["ep_rew_mean"], hist[-1]["ep_rew_mean"])


def test_recurrent_learner_runs_off_the_gpu_env():
    """The LSTM learner (the reference's RecurrentPPO, trainer.py:92-121) for ego and partner on the GPU env:
    a few iterations plus the deterministic evaluation, finite losses, episodes counted."""
    import train_ppo
    hist = train_ppo.main(["--envs", "1024", "--n-steps", "32", "--iters", "6", "--log-every", "3", "--batch-size", "8192",
                           "--max-num-timesteps", "60", "--recurrent", "--lstm-hidden", "64", "--eval-steps", "61",
                           "--device", DEV])
    assert len(hist) == 3 and hist[-1]["eval"] and hist[-1]["episodes"] >= 1024
    for h in hist[:2]:
        assert all(np.isfinite(v) for v in h["ego_loss"].values())
        assert h["episodes"] > 0 and h["partner_updates"] >= 1


def test_saved_policies_are_evaluated_by_evaluate_policy(tmp_path):
    """train_ppo --save-dir -> evaluate_policy (trainer.py:129-133 -> tester.py:64-128) on the GPU env."""
    import evaluate_policy
    import train_ppo
    d = str(tmp_path / "model")
    train_ppo.main(["--envs", "1024", "--n-steps", "32", "--iters", "4", "--log-every", "4", "--batch-size", "8192",
                    "--max-num-timesteps", "50", "--device", DEV, "--save-dir", d])
    out = evaluate_policy.main(["--max-num-timesteps", "50", "--ego-load", d + "/ppo_ego.pt", "--alt-load", d + "/ppo_partner1.pt",
                            "-t", "3000", "--envs", "1024", "-d", DEV])
    assert out["episodes"] == 3000 and out["ep_len_mean"] <= 50
    assert np.isfinite(out["average_reward"]) and out["standard_deviation"] > 0.0


def test_graphed_training_matches_its_own_eager_bookkeeping():
    """CUDA-graphed rollout + graphed minibatch updates (train_ppo's default on a GPU): the step kernel writes the
    observations into the rollout buffers' own storage, episode statistics keep counting across replays, the partner
    keeps training between replays, and learning still makes progress; `--no-graph` runs the same loop eagerly."""
    import train_ppo
    common = ["--envs", "4096", "--n-steps", "32", "--iters", "40", "--log-every", "10", "--batch-size", "32768",
              "--max-num-timesteps", "100", "--device", DEV]
    learners = []
    hist = train_ppo.main(common, learners_out=learners)
    assert all(h["cuda_graphs"] for h in hist) and len(hist) == 4
    for h in hist:
        assert all(np.isfinite(v) for v in h["ego_loss"].values()) and h["episodes"] > 0
    assert hist[-1]["partner_updates"] >= 38
    assert hist[-1]["ep_rew_mean"] > hist[0]["ep_rew_mean"]
    ego, partner = learners
    # the rollout buffers alias the env's observation ring: same memory, no copies
    assert ego.buffer.obs.data_ptr() != partner.buffer.obs.data_ptr()
    assert ego.buffer.obs.stride() == partner.buffer.obs.stride() and not ego.buffer.obs.is_contiguous()
    eager = train_ppo.main(common + ["--no-graph"])
    assert not eager[0]["cuda_graphs"] and eager[-1]["ep_rew_mean"] > eager[0]["ep_rew_mean"]
    # same environment, same learner, same schedule: the two runs see episodes of the same length budget
    assert abs(eager[-1]["episodes"] - hist[-1]["episodes"]) <= 0.2 * hist[-1]["episodes"]


def test_sb3_vecenv_adapter():
    from gym_comm_b200.pantheon import SB3VecEnvAdapter
    E, T = 16, 12
    rng = np.random.default_rng(0)
    table = torch.zeros((64, E, 2), dtype=torch.int32, device=DEV)
    table[..., 0] = torch.from_numpy(rng.integers(0, 4, (64, E))).to(DEV)
    venv = SB3VecEnvAdapter(PantheonVecEnv(OvercookedVecEnv(_ns(T), num_envs=E, device=DEV, seed=2), ScriptedPartner(table)),
                            dict_obs=True)
    obs = venv.reset()
    assert set(obs.keys()) == set(venv.observation_space.spaces.keys()) and obs["timestep"].shape == (E, 1)
    assert venv.action_space.nvec.tolist() == [4, 10] and venv.num_envs == E
    seen_terminal = 0
    prev = obs
    for t in range(30):
        acts = np.stack([rng.integers(0, 4, E), rng.integers(0, 10, E)], -1)
        venv.step_async(acts)
        obs, rew, done, infos = venv.step_wait()
        assert rew.dtype == np.float32 and rew.shape == (E,) and done.dtype == bool and len(infos) == E
        if (t + 1) % T == 0:
            assert done.all()
            for i in range(E):
                to = infos[i]["terminal_observation"]
                assert np.array_equal(to["agent1_location"], prev["agent1_location"][i])   # previous ego obs
                seen_terminal += 1
            assert np.all(obs["timestep"] == 0)                    # first obs of the new episode
        else:
            assert not done.any() and all(info == {} for info in infos)
        prev = obs
    assert seen_terminal == 2 * E
    assert venv.env_is_wrapped(object) == [False] * E and len(venv.env_method("render", indices=[0, 1])) == 2
    venv.close()
