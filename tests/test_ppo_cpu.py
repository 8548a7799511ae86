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


def test_recurrent_rollout_replay_reproduces_the_recorded_log_probs():
    """Before any update the PPO ratio must be exactly 1: replaying the stored rollout from the
    snapshotted LSTM state (with the same episode-start resets) gives the recorded values / log-probs."""
    from gym_comm_b200.ppo import RecurrentPPO
    torch.manual_seed(0)
    E, T, F = 12, 9, 5
    cfg = PPOConfig(n_steps=T, batch_size=10 ** 6)
    m = RecurrentPPO(F, 4, 3, E, "cpu", cfg, seed=2, lstm_hidden=16)
    gen = torch.Generator().manual_seed(1)
    starts = torch.ones(E)
    for rollout in range(2):                    # the second rollout starts from a non-zero carried state
        m.buffer.reset()
        for t in range(T):
            obs = torch.randn(E, F, generator=gen)
            a, v, lp = m.act(obs, starts)
            m.buffer.add(obs, a, starts, v, lp)
            starts = (torch.rand(E, generator=gen) < 0.25).float()
        with torch.no_grad():
            values, logp, _ = m.evaluate_rollout(torch.arange(E))
        assert torch.allclose(logp, m.buffer.log_probs, atol=1e-5)
        assert torch.allclose(values, m.buffer.values, atol=1e-5)
    assert m.state[0].abs().sum() > 0


def test_recurrent_ppo_learns_a_task_that_needs_memory():
    """A cue (0 / 1) is visible only at the first step of a 3-step episode; reward 1 at the last step
    iff nav == cue.  The observation at the deciding step carries no information, so only a policy
    with memory gets above chance."""
    from gym_comm_b200.ppo import RecurrentPPO
    torch.manual_seed(0)
    E, L = 256, 3
    cfg = PPOConfig(n_steps=2 * L, batch_size=E * 2 * L, n_epochs=4, learning_rate=3e-3, clip_range=0.2, ent_coef=0.0,
                    gamma=0.99)
    agent = BatchedOnPolicyAgent(RecurrentPPO(3, 2, 2, E, "cpu", cfg, seed=3, lstm_hidden=32))
    gen = torch.Generator().manual_seed(0)
    cue = torch.randint(0, 2, (E,), generator=gen)
    t_in_ep, hist = 0, []
    for it in range(1500):
        obs = torch.zeros(E, 3)
        obs[:, 2] = 1.0 if t_in_ep == 0 else 0.0                         # "a cue is shown now"
        if t_in_ep == 0:
            obs[torch.arange(E), cue] = 1.0
        a = agent.get_action(obs)
        last = t_in_ep == L - 1
        r = (a[:, 0] == cue).float() if last else torch.zeros(E)
        agent.update(r, torch.full((E,), float(last)))
        if last:
            hist.append(r.mean().item())
            cue = torch.randint(0, 2, (E,), generator=gen)
        t_in_ep = (t_in_ep + 1) % L
    assert sum(hist[:10]) / 10 < 0.65
    assert sum(hist[-10:]) / 10 > 0.9, sum(hist[-10:]) / 10
