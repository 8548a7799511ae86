"""Host-side learner logic on CPU: GAE against a hand computation, and PPO actually learning a
trivial contextual task through the same buffer / agent classes the GPU trainer uses."""
import torch

from gym_comm_b200.ppo import PPO, PPOConfig, RolloutBuffer
from gym_comm_b200.pantheon import BatchedOnPolicyAgent


def test_gae_matches_hand_computation():
    buf = RolloutBuffer(3, 1, 2, "cpu", gamma=0.5, gae_lambda=1.0)
    vals = [1.0, 2.0, 3.0]
    rews = [1.0, 0.0, 2.0]
    starts = [1.0, 0.0, 1.0]          # a new episode starts at step 2
    for i in range(3):
        buf.add(torch.zeros(1, 2), torch.zeros(1, 2, dtype=torch.int64), torch.tensor([starts[i]]),
                torch.tensor([vals[i]]), torch.zeros(1))
        buf.add_reward(torch.tensor([rews[i]]))
    buf.compute_returns_and_advantage(torch.tensor([4.0]), torch.tensor([0.0]))
    # step 2: delta = 2 + .5*4 - 3 = 1 ; step 1: next is an episode start -> delta = 0 - 2 = -2 ;
    # step 0: delta = 1 + .5*2 - 1 = 1, gae = 1 + .5*(-2) = 0
    assert torch.allclose(buf.advantages[:, 0], torch.tensor([0.0, -2.0, 1.0]))
    assert torch.allclose(buf.returns[:, 0], torch.tensor([1.0, 0.0, 4.0]))


def test_ppo_learns_contextual_bandit():
    """obs one-hot of 4 contexts; reward 1 iff nav == context and comm == 3 - context."""
    torch.manual_seed(0)
    E, C = 256, 4
    cfg = PPOConfig(n_steps=8, batch_size=1024, n_epochs=4, learning_rate=3e-3, clip_range=0.2, ent_coef=0.0, gamma=0.0)
    agent = BatchedOnPolicyAgent(PPO(4, 4, C, E, "cpu", cfg, seed=1))
    gen = torch.Generator().manual_seed(0)
    mean_rew = []
    ctx = torch.randint(0, 4, (E,), generator=gen)
    for it in range(400):
        obs = torch.nn.functional.one_hot(ctx, 4).float()
        a = agent.get_action(obs)
        r = ((a[:, 0] == ctx) & (a[:, 1] == 3 - ctx)).float()
        agent.update(r, torch.ones(E))
        mean_rew.append(r.mean().item())
        ctx = torch.randint(0, 4, (E,), generator=gen)
    assert sum(mean_rew[:20]) / 20 < 0.2
    assert sum(mean_rew[-20:]) / 20 > 0.8, sum(mean_rew[-20:]) / 20
    assert agent.iteration >= 40
