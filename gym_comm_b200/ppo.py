"""A small batched PPO (torch) for the MultiDiscrete([4, C]) action space of the Overcooked env.

The reference trains with SB3 / sb3_contrib (`trainer.py:92-121`), which is not installed here
and is out of scope to rebuild; this module is the minimum learner needed to show the GPU env
training end to end.  It follows SB3's PPO where that matters for
comparability: MlpPolicy-style separate pi / vf towers (2 x 64, tanh, orthogonal init), GAE with
episode-start masks (`RolloutBuffer.compute_returns_and_advantage`), per-minibatch advantage
normalisation, clipped surrogate, unclipped value loss, entropy bonus, grad-norm clipping.
Observations are the flat float32 rows the kernel writes == what `FlattenedDictExtractor`
(gym_comm/extractors/CustomExtractor.py:119-128) would concatenate.  Everything stays on the
env's device; nothing syncs to the host inside a rollout.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn as nn


def _mlp(inp: int, hidden: Tuple[int, ...]) -> nn.Sequential:
    layers, d = [], inp
    for h in hidden:
        lin = nn.Linear(d, h)
        nn.init.orthogonal_(lin.weight, gain=math.sqrt(2))
        nn.init.zeros_(lin.bias)
        layers += [lin, nn.Tanh()]
        d = h
    return nn.Sequential(*layers)


# ---- data-parallel learners ("the only collective in the whole system would be an NCCL all-reduce
# of PPO gradients").  One process per GPU, each with its own env shard and rollout buffer; the policies are a few
# hundred KB, so every minibatch ends in ONE all-reduce over a flat gradient buffer.  No-ops without a process group.
def _world() -> int:
    import torch.distributed as dist
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def broadcast_parameters(module: nn.Module, src: int = 0) -> None:
    """Every rank starts from rank `src`'s weights."""
    if _world() == 1:
        return
    import torch.distributed as dist
    with torch.no_grad():
        flat = torch.cat([p.reshape(-1) for p in module.parameters()])
        dist.broadcast(flat, src)
        o = 0
        for p in module.parameters():
            p.copy_(flat[o:o + p.numel()].view_as(p))
            o += p.numel()


def average_gradients(module: nn.Module) -> None:
    """Mean of the ranks' gradients, in place (call between backward() and the optimizer step)."""
    w = _world()
    if w == 1:
        return
    import torch.distributed as dist
    ps = list(module.parameters())
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in ps])
    dist.all_reduce(flat)
    flat /= w
    o = 0
    for p in ps:
        g = flat[o:o + p.numel()].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        o += p.numel()


class Cat:
    """Categorical over the last axis of `logits` with the four things PPO needs.  No argument validation and no host
    synchronisation (torch.distributions checks its inputs with `.all()`), so it can sit inside a CUDA graph; sampling
    is Gumbel-max over torch's own (graph-safe) Philox stream."""

    def __init__(self, logits):
        self.logp = torch.log_softmax(logits, -1)

    @property
    def probs(self):
        return self.logp.exp()

    def sample(self):
        u = torch.rand_like(self.logp)
        return (self.logp - torch.log(-torch.log(u))).argmax(-1)

    def log_prob(self, a):
        return self.logp.gather(-1, a.unsqueeze(-1)).squeeze(-1)

    def entropy(self):
        return -(self.logp.exp() * self.logp).sum(-1)


class ActorCritic(nn.Module):
    """obs [.., F] -> logits over nav (4) and message (C), value."""

    def __init__(self, obs_dim: int, num_nav: int, num_comm: int, hidden=(64, 64)):
        super().__init__()
        self.num_nav, self.num_comm = num_nav, num_comm
        self.pi = _mlp(obs_dim, hidden)
        self.vf = _mlp(obs_dim, hidden)
        self.action_head = nn.Linear(hidden[-1], num_nav + num_comm)
        self.value_head = nn.Linear(hidden[-1], 1)
        nn.init.orthogonal_(self.action_head.weight, gain=0.01)
        nn.init.zeros_(self.action_head.bias)
        nn.init.orthogonal_(self.value_head.weight, gain=1.0)
        nn.init.zeros_(self.value_head.bias)

    def _dists(self, obs):
        logits = self.action_head(self.pi(obs))
        return Cat(logits[..., :self.num_nav]), Cat(logits[..., self.num_nav:])

    def value(self, obs):
        return self.value_head(self.vf(obs)).squeeze(-1)

    @torch.no_grad()
    def act(self, obs, deterministic: bool = False):
        """-> actions int64 [.., 2], values [..], log_probs [..]   (SB3 `policy.forward`)."""
        dn, dc = self._dists(obs)
        nav = dn.probs.argmax(-1) if deterministic else dn.sample()
        com = dc.probs.argmax(-1) if deterministic else dc.sample()
        return torch.stack([nav, com], -1), self.value(obs), dn.log_prob(nav) + dc.log_prob(com)

    def evaluate(self, obs, actions):
        """-> values, log_probs, entropy   (SB3 `policy.evaluate_actions`)."""
        dn, dc = self._dists(obs)
        logp = dn.log_prob(actions[..., 0]) + dc.log_prob(actions[..., 1])
        return self.value(obs), logp, dn.entropy() + dc.entropy()


class RolloutBuffer:
    """[n_steps, E, ...] tensors on the device; SB3 `RolloutBuffer` semantics."""

    def __init__(self, n_steps: int, num_envs: int, obs_dim: int, device, gamma=0.99, gae_lambda=0.95):
        self.n_steps, self.num_envs, self.gamma, self.gae_lambda = n_steps, num_envs, gamma, gae_lambda
        kw = dict(device=device)
        # `obs` may be re-pointed at a view of storage the ENV writes (PantheonVecEnv.attach_rollout_storage): the step
        # kernel then fills the rollout buffer itself and `add` has nothing to copy
        self.obs = torch.zeros((n_steps, num_envs, obs_dim), **kw)
        self.actions = torch.zeros((n_steps, num_envs, 2), dtype=torch.int64, **kw)
        self.rewards = torch.zeros((n_steps, num_envs), **kw)
        self.episode_starts = torch.zeros((n_steps, num_envs), **kw)
        self.values = torch.zeros((n_steps, num_envs), **kw)
        self.log_probs = torch.zeros((n_steps, num_envs), **kw)
        self.advantages = torch.zeros((n_steps, num_envs), **kw)
        self.returns = torch.zeros((n_steps, num_envs), **kw)
        self.pos = 0

    @property
    def full(self) -> bool:
        return self.pos >= self.n_steps

    def reset(self):
        self.pos = 0
        self.rewards.zero_()

    def add(self, obs, actions, episode_starts, values, log_probs):
        p = self.pos
        if obs.data_ptr() != self.obs[p].data_ptr():       # already in place when the env wrote this slot itself
            self.obs[p].copy_(obs)
        self.actions[p].copy_(actions)
        self.episode_starts[p].copy_(episode_starts)
        self.values[p].copy_(values)
        self.log_probs[p].copy_(log_probs)
        self.pos += 1

    def add_reward(self, reward):
        """Reward of the most recently recorded action (OnPolicyAgent.update, agents.py:196-213)."""
        self.rewards[self.pos - 1] += reward

    def compute_returns_and_advantage(self, last_values, dones):
        last_gae = torch.zeros_like(last_values)
        for step in reversed(range(self.n_steps)):
            if step == self.n_steps - 1:
                next_non_terminal, next_values = 1.0 - dones, last_values
            else:
                next_non_terminal, next_values = 1.0 - self.episode_starts[step + 1], self.values[step + 1]
            delta = self.rewards[step] + self.gamma * next_values * next_non_terminal - self.values[step]
            last_gae = delta + self.gamma * self.gae_lambda * next_non_terminal * last_gae
            self.advantages[step] = last_gae
        torch.add(self.advantages, self.values, out=self.returns)


@dataclass
class PPOConfig:
    """Defaults = the reference's (`trainer.py:92-112`) except n_steps, which there is per single env
    (5000) and here is per env of a large batch."""
    n_steps: int = 128
    batch_size: int = 16384
    n_epochs: int = 4
    learning_rate: float = 3e-4
    gamma: float = 0.99
    gae_lambda: float = 0.95
    clip_range: float = 0.05
    ent_coef: float = 0.01
    vf_coef: float = 0.5
    max_grad_norm: float = 0.5
    cuda_graph: bool = True      # capture the minibatch update (gather, forward, backward, [all-reduce,] clip, Adam) in a
                                 # CUDA graph when the learner sits on a GPU; eager otherwise

    @staticmethod
    def from_hyperparams(h: dict, **over):
        """`hyperparams` block of the reference's JSON configs (n_steps, batch_size, learning_rate,
        entrop_coef, clip_range)."""
        c = PPOConfig()
        c.learning_rate = float(h.get("learning_rate", c.learning_rate))
        c.ent_coef = float(h.get("entrop_coef", c.ent_coef))
        c.clip_range = float(h.get("clip_range", c.clip_range))
        for k, v in over.items():
            setattr(c, k, v)
        return c


class PPO:
    def __init__(self, obs_dim: int, num_nav: int, num_comm: int, num_envs: int, device, cfg: PPOConfig = None,
                 seed: int = 0):
        self.cfg = cfg or PPOConfig()
        self.device = torch.device(device)
        gen = torch.Generator().manual_seed(seed)
        with torch.random.fork_rng(devices=[]):
            torch.manual_seed(int(torch.randint(0, 2 ** 31 - 1, (1,), generator=gen)))
            self.policy = ActorCritic(obs_dim, num_nav, num_comm).to(self.device)
        broadcast_parameters(self.policy)
        on_gpu = self.device.type == "cuda"
        self.optimizer = torch.optim.Adam(self.policy.parameters(), lr=self.cfg.learning_rate, eps=1e-5,
                                          **(dict(fused=True, capturable=True) if on_gpu else {}))
        self.buffer = RolloutBuffer(self.cfg.n_steps, num_envs, obs_dim, self.device, self.cfg.gamma, self.cfg.gae_lambda)
        self.n_updates = 0
        self._graph = None           # (CUDAGraph, static index buffer, static stats) of the captured minibatch update

    # uniform entry points of the rollout code (the recurrent learner keeps LSTM state behind them)
    def act(self, obs, episode_starts=None, deterministic: bool = False):
        return self.policy.act(obs, deterministic)

    def value(self, obs, episode_starts=None):
        with torch.no_grad():
            return self.policy.value(obs)

    def _minibatch(self, obs, actions, old_logp, adv_all, ret, idx, acc):
        """One clipped-surrogate update on the samples `idx`; the five statistics are ADDED to `acc` on the device (no
        host synchronisation, so the call can be captured in a CUDA graph)."""
        c = self.cfg
        values, logp, ent = self.policy.evaluate(obs[idx], actions[idx])
        adv = adv_all[idx]
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        olp = old_logp[idx]
        ratio = torch.exp(logp - olp)
        pg = -torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - c.clip_range, 1 + c.clip_range)).mean()
        vf = torch.nn.functional.mse_loss(values, ret[idx])
        entm = ent.mean()
        loss = pg + c.vf_coef * vf - c.ent_coef * entm
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        average_gradients(self.policy)
        nn.utils.clip_grad_norm_(self.policy.parameters(), c.max_grad_norm)
        self.optimizer.step()
        with torch.no_grad():
            acc += torch.stack([pg.detach(), vf.detach(), entm.detach(),
                                ((ratio - 1).abs() > c.clip_range).float().mean(), (olp - logp).mean().detach()])

    def train(self):
        """One PPO update over the (full) rollout buffer.  Returns a dict of float stats (ONE host read at the end)."""
        c, b = self.cfg, self.buffer
        n = b.n_steps * b.num_envs
        if b.obs.is_contiguous():
            obs = b.obs.view(n, -1)
        else:                                       # a strided view of the env's own storage: one gather per update, into a
            if getattr(self, "_obs_flat", None) is None or self._obs_flat.shape[0] != n:    # buffer that keeps its address
                self._obs_flat = torch.empty((n, b.obs.shape[-1]), device=self.device)
            self._obs_flat.view(b.obs.shape).copy_(b.obs)
            obs = self._obs_flat
        actions = b.actions.reshape(n, 2)
        old_logp, adv_all, ret = b.log_probs.reshape(n), b.advantages.reshape(n), b.returns.reshape(n)
        bs = min(c.batch_size, n)
        acc = torch.zeros(5, device=self.device)
        nmb = 0
        # The first update runs eagerly (it is also the warm-up every capture needs); from the second one on, the whole
        # minibatch update -- gather, forward, backward, gradient clipping, the (capturable, fused) Adam step -- is ONE
        # CUDA graph over static tensors, replayed per minibatch with a fresh index vector.
        # (Data parallel: the gradient all-reduce is part of the captured update -- NCCL collectives can be captured;
        # `OC_PPO_GRAPH_DP=0` keeps the multi-rank update eager.)
        graphed = (c.cuda_graph and self.device.type == "cuda" and self.n_updates > 0 and
                   (_world() == 1 or os.environ.get("OC_PPO_GRAPH_DP", "1") != "0"))
        if graphed:
            g = self._graph
            key = (n, bs, obs.data_ptr(), actions.data_ptr(), old_logp.data_ptr(), adv_all.data_ptr(), ret.data_ptr())
            if g is None or g["key"] != key:
                idx_buf = torch.zeros(bs, dtype=torch.int64, device=self.device)
                acc_buf = torch.zeros(5, device=self.device)
                graph = torch.cuda.CUDAGraph()
                self.optimizer.zero_grad(set_to_none=True)
                torch.cuda.synchronize(self.device)
                with torch.cuda.graph(graph):
                    self._minibatch(obs, actions, old_logp, adv_all, ret, idx_buf, acc_buf)
                g = self._graph = dict(key=key, graph=graph, idx=idx_buf, acc=acc_buf)
            g["acc"].zero_()
        for _ in range(c.n_epochs):
            perm = torch.randperm(n, device=self.device)
            for i in range(0, n - bs + 1, bs):
                if graphed:
                    g["idx"].copy_(perm[i:i + bs])
                    g["graph"].replay()
                else:
                    self._minibatch(obs, actions, old_logp, adv_all, ret, perm[i:i + bs], acc)
                nmb += 1
        self.n_updates += 1
        vals = ((g["acc"] if graphed else acc) / max(nmb, 1)).tolist()
        return dict(zip(("pg", "vf", "ent", "clipfrac", "kl"), vals))

# ------------------------------------------------------------------------------------------------
# Recurrent learner: what the reference actually trains (`RecurrentPPO("MultiInputPolicy", ...)`,
# trainer.py:92-121; sb3_contrib/ppo_recurrent): separate LSTMs for actor and critic in front of the
# 2 x 64 towers, hidden state reset at episode starts, back-propagation through the whole rollout.
# ------------------------------------------------------------------------------------------------
class RecurrentActorCritic(nn.Module):
    """obs [.., F] + LSTM state -> logits over nav (4) and message (C), value.  State = (h_pi, c_pi,
    h_vf, c_vf), each [E, H] (one layer, like sb3_contrib's default n_lstm_layers=1)."""

    def __init__(self, obs_dim: int, num_nav: int, num_comm: int, lstm_hidden: int = 256, hidden=(64, 64)):
        super().__init__()
        self.num_nav, self.num_comm, self.lstm_hidden = num_nav, num_comm, lstm_hidden
        self.lstm_pi = nn.LSTMCell(obs_dim, lstm_hidden)
        self.lstm_vf = nn.LSTMCell(obs_dim, lstm_hidden)
        self.pi = _mlp(lstm_hidden, hidden)
        self.vf = _mlp(lstm_hidden, hidden)
        self.action_head = nn.Linear(hidden[-1], num_nav + num_comm)
        self.value_head = nn.Linear(hidden[-1], 1)
        nn.init.orthogonal_(self.action_head.weight, gain=0.01)
        nn.init.zeros_(self.action_head.bias)
        nn.init.orthogonal_(self.value_head.weight, gain=1.0)
        nn.init.zeros_(self.value_head.bias)

    def initial_state(self, num_envs: int, device):
        z = torch.zeros((num_envs, self.lstm_hidden), device=device)
        return (z, z.clone(), z.clone(), z.clone())

    def step(self, obs, state, episode_starts):
        """One time step for [E] envs: the state of envs that start an episode is zeroed first
        (sb3_contrib `_process_sequence`).  -> (dist_nav, dist_comm, value, new_state)"""
        keep = (1.0 - episode_starts).unsqueeze(-1)
        hp, cp = self.lstm_pi(obs, (state[0] * keep, state[1] * keep))
        hv, cv = self.lstm_vf(obs, (state[2] * keep, state[3] * keep))
        logits = self.action_head(self.pi(hp))
        value = self.value_head(self.vf(hv)).squeeze(-1)
        return Cat(logits[..., :self.num_nav]), Cat(logits[..., self.num_nav:]), value, (hp, cp, hv, cv)


class RecurrentPPO:
    """PPO over `RecurrentActorCritic`; same rollout-side interface as `PPO` (`act`, `value`, `buffer`,
    `train`).  The LSTM state lives here, one row per env, and is snapshotted when a rollout starts so
    that `train` can replay the rollout from the same state."""

    def __init__(self, obs_dim: int, num_nav: int, num_comm: int, num_envs: int, device, cfg: PPOConfig = None,
                 seed: int = 0, lstm_hidden: int = 256):
        self.cfg = cfg or PPOConfig()
        self.device = torch.device(device)
        gen = torch.Generator().manual_seed(seed)
        with torch.random.fork_rng(devices=[]):
            torch.manual_seed(int(torch.randint(0, 2 ** 31 - 1, (1,), generator=gen)))
            self.policy = RecurrentActorCritic(obs_dim, num_nav, num_comm, lstm_hidden).to(self.device)
        broadcast_parameters(self.policy)
        self.optimizer = torch.optim.Adam(self.policy.parameters(), lr=self.cfg.learning_rate, eps=1e-5)
        self.buffer = RolloutBuffer(self.cfg.n_steps, num_envs, obs_dim, self.device, self.cfg.gamma, self.cfg.gae_lambda)
        self.state = self.policy.initial_state(num_envs, self.device)
        self.rollout_state = self.state
        self.n_updates = 0

    @torch.no_grad()
    def act(self, obs, episode_starts, deterministic: bool = False):
        if self.buffer.pos == 0:                       # first step of a rollout: remember where it started
            self.rollout_state = tuple(s.clone() for s in self.state)
        dn, dc, value, self.state = self.policy.step(obs, self.state, episode_starts)
        nav = dn.probs.argmax(-1) if deterministic else dn.sample()
        com = dc.probs.argmax(-1) if deterministic else dc.sample()
        return torch.stack([nav, com], -1), value, dn.log_prob(nav) + dc.log_prob(com)

    @torch.no_grad()
    def value(self, obs, episode_starts):
        """Value of `obs` as the NEXT step would see it; the state is not advanced."""
        return self.policy.step(obs, self.state, episode_starts)[2]

    def evaluate_rollout(self, env_idx):
        """Replay the stored rollout of the envs `env_idx` from its initial state.
        -> values, log_probs, entropy, each [n_steps, len(env_idx)]"""
        b = self.buffer
        state = tuple(s[env_idx] for s in self.rollout_state)
        vals, logps, ents = [], [], []
        for t in range(b.n_steps):
            dn, dc, v, state = self.policy.step(b.obs[t, env_idx], state, b.episode_starts[t, env_idx])
            a = b.actions[t, env_idx]
            vals.append(v)
            logps.append(dn.log_prob(a[..., 0]) + dc.log_prob(a[..., 1]))
            ents.append(dn.entropy() + dc.entropy())
        return torch.stack(vals), torch.stack(logps), torch.stack(ents)

    def train(self):
        """One PPO update; minibatches are sets of ENVS (whole sequences), back-propagated through time."""
        c, b = self.cfg, self.buffer
        E, T = b.num_envs, b.n_steps
        envs_per_batch = max(1, min(E, c.batch_size // T))
        acc = torch.zeros(5, device=self.device)             # statistics stay on the device until the update is over
        nmb = 0
        for _ in range(c.n_epochs):
            perm = torch.randperm(E, device=self.device)
            for i in range(0, E - envs_per_batch + 1, envs_per_batch):
                idx = perm[i:i + envs_per_batch]
                values, logp, ent = self.evaluate_rollout(idx)
                adv = b.advantages[:, idx]
                adv = (adv - adv.mean()) / (adv.std() + 1e-8)
                old_logp = b.log_probs[:, idx]
                ratio = torch.exp(logp - old_logp)
                pg = -torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - c.clip_range, 1 + c.clip_range)).mean()
                vf = torch.nn.functional.mse_loss(values, b.returns[:, idx])
                entm = ent.mean()
                loss = pg + c.vf_coef * vf - c.ent_coef * entm
                self.optimizer.zero_grad(set_to_none=True)
                loss.backward()
                average_gradients(self.policy)
                nn.utils.clip_grad_norm_(self.policy.parameters(), c.max_grad_norm)
                self.optimizer.step()
                with torch.no_grad():
                    acc += torch.stack([pg.detach(), vf.detach(), entm.detach(),
                                        ((ratio - 1).abs() > c.clip_range).float().mean(), (old_logp - logp).mean().detach()])
                nmb += 1
        self.n_updates += 1
        return dict(zip(("pg", "vf", "ent", "clipfrac", "kl"), (acc / max(nmb, 1)).tolist()))


# ------------------------------------------------------------------------------------------------
# Saving / loading a learner's policy -- the counterpart of `ego.save(path)` / `PPO.load(path)`
# (trainer.py:129-133, tester.py:64-70).  One file per agent, self-describing (kind + sizes).
# ------------------------------------------------------------------------------------------------
def save_learner(learner, path: str) -> None:
    pol = learner.policy
    recurrent = isinstance(pol, RecurrentActorCritic)
    first = pol.lstm_pi.weight_ih if recurrent else pol.pi[0].weight
    meta = dict(kind="lstm" if recurrent else "mlp", obs_dim=int(first.shape[1]), num_nav=int(pol.num_nav),
                num_comm=int(pol.num_comm), lstm_hidden=int(pol.lstm_hidden) if recurrent else 0)
    torch.save({"meta": meta, "policy": pol.state_dict()}, path)


def load_learner(path: str, num_envs: int, device, cfg: PPOConfig = None):
    """-> PPO or RecurrentPPO for `num_envs` envs with the saved weights (optimizer state is not kept)."""
    blob = torch.load(path, map_location=device, weights_only=True)
    m = blob["meta"]
    if m["kind"] == "lstm":
        learner = RecurrentPPO(m["obs_dim"], m["num_nav"], m["num_comm"], num_envs, device, cfg, lstm_hidden=m["lstm_hidden"])
    else:
        learner = PPO(m["obs_dim"], m["num_nav"], m["num_comm"], num_envs, device, cfg)
    learner.policy.load_state_dict(blob["policy"])
    return learner
