"""Batched PantheonRL layer: the ego-centric view of the 2-player env with the partner learner
living INSIDE `step`, re-expressed over [E] tensors.

Reference semantics mirrored (pantheonrl/common/multiagentenv.py:149-243, agents.py:112-213):
* `step(ego_action)`: the partner's action comes from `partner.get_action(partner_obs)` on the
  observation the partner saw after the previous step; both act simultaneously
  (`SimultaneousEnv.n_step`, :395-404); the partner then receives `update(reward, done)`.
* the ego's reward is the env reward of this step; on `done` the reference returns the PREVIOUS
  ego observation (:206-208) -- which SB3's DummyVecEnv turns into
  `infos["terminal_observation"]` before resetting.  Here: `terminal_obs` holds that previous
  observation for the envs that finished, and the returned observation is the first one of the
  new episode (VecEnv auto-reset contract).
* `OnPolicyAgent`: records (obs, action, value, log_prob, episode_start) at `get_action`, adds
  the reward to the last recorded transition at `update`, and trains when its buffer is full --
  inside the env step, exactly where the reference does it.
"""
from __future__ import annotations

from typing import Optional

import torch

from .ppo import PPO
from .vec_env import OvercookedVecEnv


class BatchedOnPolicyAgent:
    """`OnPolicyAgent` (pantheonrl/common/agents.py:84-213) over E envs at once.  Its per-env state lives in tensors
    that keep their address (updated in place), so a whole rollout can be captured in a CUDA graph (`GraphedRollout`)."""

    def __init__(self, model: PPO, name: str = "partner"):
        self.model = model
        self.name = name
        self._last_episode_starts = torch.ones(model.buffer.num_envs, device=model.device)
        self._values = torch.zeros(model.buffer.num_envs, device=model.device)
        self.num_timesteps = 0
        self.iteration = 0
        self.last_train_stats = None

    def maybe_train(self) -> bool:
        """Train the model if the buffer is full (agents.py:127-160) -- the reference does this at the top of
        `get_action`, i.e. inside the env step that follows the one which filled the buffer."""
        buf = self.model.buffer
        if not buf.full:
            return False
        buf.compute_returns_and_advantage(self._values, self._last_episode_starts)
        self.last_train_stats = self.model.train()
        self.iteration += 1
        buf.reset()
        return True

    def get_action(self, obs: torch.Tensor, record: bool = True) -> torch.Tensor:
        buf = self.model.buffer
        if record:
            self.maybe_train()
        actions, values, log_probs = self.model.act(obs, self._last_episode_starts)
        if record:
            buf.add(obs, actions, self._last_episode_starts, values, log_probs)
        self.num_timesteps += obs.shape[0]
        self._values.copy_(values)
        return actions

    def update(self, reward: torch.Tensor, done: torch.Tensor) -> None:
        self._last_episode_starts.copy_(done)
        self.model.buffer.add_reward(reward)


class BatchedStaticPolicyAgent:
    """`StaticPolicyAgent` (pantheonrl/common/agents.py:55-80) over E envs: acts from a fixed policy, never
    learns.  Like the reference it SAMPLES from the policy (`action_from_policy` calls `policy.forward`);
    `deterministic=True` takes the argmax instead.  `update` only tracks episode starts, which a recurrent
    policy needs to reset its LSTM state."""

    def __init__(self, model, deterministic: bool = False):
        self.model = model
        self.deterministic = bool(deterministic)
        self._starts = torch.ones(model.buffer.num_envs, device=model.device)

    def get_action(self, obs: torch.Tensor, record: bool = True) -> torch.Tensor:
        return self.model.act(obs, self._starts, deterministic=self.deterministic)[0]

    def update(self, reward: torch.Tensor, done: torch.Tensor) -> None:
        self._starts.copy_(done)


class PantheonVecEnv:
    """Ego-centric batched env: `reset() -> ego_obs [E, F]`, `step(ego_actions [E, 2]) ->
    (ego_obs, reward [E], done [E] u8)`; the partner acts and learns inside `step`.

    Observations live in a ring of `[slots, E, A, F]` rows the step kernel writes straight into (`obs_out`): the row
    of step t is never copied again -- the partner reads its half in place, and after `attach_rollout_storage` the
    rollout buffers of the ego and the partner learner ARE views of that ring."""

    def __init__(self, env: OvercookedVecEnv, partner: Optional[BatchedOnPolicyAgent] = None, ego_ind: int = 0,
                 reward_scale: float = 1.0):
        if env.num_agents != 2:
            raise ValueError("the PantheonRL layer is 2-player (SimultaneousEnv, multiagentenv.py:390-393)")
        if ego_ind != 0:
            raise ValueError("ego_ind must be 0 (as in the reference trainer)")
        if not env.auto_reset:
            raise ValueError("PantheonVecEnv needs an auto-resetting OvercookedVecEnv")
        self.env = env
        self.partner = partner
        self.reward_scale = float(reward_scale)      # learner-side scaling; episode statistics stay in env units
        self.num_envs = env.num_envs
        self.obs_dim = env.obs_width
        self.device = env.device
        E, F = self.num_envs, self.obs_dim
        self._actions = torch.zeros((E, 2, 2), dtype=torch.int32, device=self.device)
        self._ring = torch.zeros((2, E, 2, F), device=self.device)        # slot p: what the agents see before step p
        self._slot = 0
        self._rew = torch.zeros((E, 2), device=self.device)
        self._done = torch.zeros((E,), dtype=torch.uint8, device=self.device)
        self.terminal_obs = torch.zeros((E, F), device=self.device)
        self.ep_return = torch.zeros(E, device=self.device)
        self.ep_length = torch.zeros(E, device=self.device)
        # running episode statistics (device-side, read by the trainer when it logs)
        self.finished_episodes = torch.zeros((), device=self.device)
        self.finished_return_sum = torch.zeros((), device=self.device)
        self.finished_length_sum = torch.zeros((), device=self.device)
        self.finished_success = torch.zeros((), device=self.device)
        self.finished_return_sq = torch.zeros((), device=self.device, dtype=torch.float64)

    @property
    def _obs(self) -> torch.Tensor:
        """[E, 2, F]: the observations both agents currently see."""
        return self._ring[self._slot]

    def add_partner_agent(self, agent: BatchedOnPolicyAgent):
        self.partner = agent

    def attach_rollout_storage(self, n_steps: int, *learners) -> None:
        """Make the observation ring `n_steps + 1` slots long and point the rollout buffers of `learners` (player 0,
        player 1) at it: `buffer.obs[t]` is then the view `ring[t, :, player]` the env itself wrote, and
        `RolloutBuffer.add` finds the observation already in place.  After `n_steps` steps call `rollover()`."""
        E, F = self.num_envs, self.obs_dim
        cur = self._ring[self._slot].clone()
        self._ring = torch.zeros((n_steps + 1, E, 2, F), device=self.device)
        self._ring[0].copy_(cur)
        self._slot = 0
        for player, lr in enumerate(learners):
            if lr is not None:
                assert lr.buffer.n_steps == n_steps and lr.buffer.num_envs == E
                lr.buffer.obs = self._ring[:n_steps, :, player]

    def rollover(self) -> None:
        """The last slot becomes the first of the next rollout (one [E, 2, F] copy per rollout)."""
        if self._slot != 0:
            self._ring[0].copy_(self._ring[self._slot])
            self._slot = 0

    def reset(self) -> torch.Tensor:
        self._slot = 0
        self.env.reset(obs_out=self._ring[0])
        self.ep_return.zero_()
        self.ep_length.zero_()
        return self._ring[0][:, 0]

    def step(self, ego_actions: torch.Tensor):
        assert self.partner is not None, "add_partner_agent first (multiagentenv.py:92-101)"
        cur = self._ring[self._slot]
        nxt_slot = (self._slot + 1) % self._ring.shape[0]
        partner_actions = self.partner.get_action(cur[:, 1])                 # _get_actions (:149-161)
        self._actions[:, 0].copy_(ego_actions)
        self._actions[:, 1].copy_(partner_actions)
        # n_step -> multi_step: the kernel writes the next observations into the next ring slot
        obs, rew, done = self.env.step(self._actions, obs_out=self._ring[nxt_slot], rew_out=self._rew, done_out=self._done)
        self.partner.update(rew[:, 1] * self.reward_scale, done)             # _update_players (:163-170)
        d = done.bool()
        torch.where(d[:, None], cur[:, 0], self.terminal_obs, out=self.terminal_obs)   # "old ego obs" (:206-208)
        self._slot = nxt_slot
        # episode bookkeeping (in place: the tensors keep their addresses)
        self.ep_return += rew[:, 0]
        self.ep_length += 1
        T = float(self.env.arglist.max_num_timesteps)
        self.finished_episodes += d.sum()
        self.finished_return_sum += (self.ep_return * d).sum()
        self.finished_length_sum += (self.ep_length * d).sum()
        self.finished_return_sq += (self.ep_return.double() ** 2 * d).sum()
        self.finished_success += (d & (self.ep_length < T)).sum()             # ended by delivery, not by the clock
        self.ep_return.masked_fill_(d, 0.0)
        self.ep_length.masked_fill_(d, 0.0)
        return obs[:, 0], rew[:, 0] * self.reward_scale, done

    def pop_episode_stats(self):
        """Host read (one sync) of the episode statistics accumulated since the last call; summed over the
        ranks of a data-parallel job (every rank must call it)."""
        acc = (self.finished_episodes, self.finished_return_sum, self.finished_length_sum, self.finished_success,
               self.finished_return_sq)
        # what the reference's EpisodeRecorder logs at every reset (episode_recorder.py:29): completed subtasks of the
        # last finished episode, kept per env on the device (oc_get_stats); here its mean over the envs that have
        # finished at least one episode
        st = self.env.stats()
        has = st["episodes"] > 0
        v = torch.stack([t.to(torch.float64) for t in acc] +
                        [(st["num_completed_subtasks"] * has).sum().to(torch.float64), has.sum().to(torch.float64)])
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(v)
        n, rsum, lsum, succ, rsq, csum, cn = v.tolist()
        mean = rsum / max(n, 1.0)
        out = dict(episodes=n, ep_rew_mean=mean, ep_rew_std=max(rsq / max(n, 1.0) - mean * mean, 0.0) ** 0.5,
                   ep_len_mean=lsum / max(n, 1.0), delivered_frac=succ / max(n, 1.0),
                   num_completed_subtasks=csum / max(cn, 1.0))
        for t in acc:
            t.zero_()
        return out


def collect_rollout(penv: PantheonVecEnv, ego: PPO, obs: torch.Tensor, episode_starts: torch.Tensor):
    """`collect_rollouts` (sb3_contrib/ppo_recurrent/ppo_recurrent.py:195-310) over [E] tensors without its per-step
    host synchronisations: n_steps x (ego acts and records, the partner acts / records / learns inside `penv.step`)."""
    buf = ego.buffer
    buf.reset()
    for _ in range(buf.n_steps):
        actions, values, log_probs = ego.act(obs, episode_starts)
        buf.add(obs, actions, episode_starts, values, log_probs)
        obs, rew, done = penv.step(actions.to(torch.int32))
        buf.add_reward(rew)
        episode_starts = done.to(torch.float32)
    return obs, episode_starts


def collect_and_train(penv: PantheonVecEnv, ego: PPO, obs: torch.Tensor, episode_starts: torch.Tensor):
    """One ego iteration: fill the ego buffer with n_steps of experience (the partner records and
    trains on its own schedule inside penv.step), then one PPO update.  Mirrors
    `collect_rollouts` + `train` of `OnPolicyAlgorithm.learn`
    (sb3_contrib/ppo_recurrent/ppo_recurrent.py:195-312) without the per-step host syncs."""
    obs, episode_starts = collect_rollout(penv, ego, obs, episode_starts)
    last_values = ego.value(obs, episode_starts)
    ego.buffer.compute_returns_and_advantage(last_values, episode_starts)
    stats = ego.train()
    return obs, episode_starts, stats


class GraphedRollout:
    """A whole rollout -- n_steps x [ego policy, partner policy, `oc_step`, rollout-buffer writes, episode
    bookkeeping] -- captured ONCE in a CUDA graph and replayed per iteration: a rollout step is a few hundred tiny
    launches, and replaying them costs the GPU time of the kernels instead of the host time of issuing them.

    Needs feed-forward learners (their per-env state is the rollout buffer; the LSTM learner carries Python-side
    state) and the ring storage of `PantheonVecEnv.attach_rollout_storage`, so that every tensor the graph touches
    keeps its address.  The partner's own update ("train when the buffer is full", the first thing its next
    `get_action` does in the reference) runs between replays, which is exactly where the reference runs it."""

    def __init__(self, penv: PantheonVecEnv, ego: PPO):
        if not isinstance(ego, PPO) or not isinstance(getattr(penv.partner, "model", None), PPO):
            raise TypeError("GraphedRollout needs feed-forward PPO learners for ego and partner")
        self.penv, self.ego = penv, ego
        n = ego.buffer.n_steps
        assert penv.partner.model.buffer.n_steps == n, "ego and partner must roll out the same number of steps"
        penv.attach_rollout_storage(n, ego, penv.partner.model)
        self.starts = torch.ones(penv.num_envs, device=penv.device)       # episode starts seen by the first step
        self.graph = None
        self.replays = 0

    def _body(self):
        penv, ego = self.penv, self.ego
        obs = penv._ring[0][:, 0]
        obs, starts = collect_rollout(penv, ego, obs, self.starts)
        self.starts.copy_(starts)
        penv.rollover()
        return penv._ring[0][:, 0]

    def run(self):
        """One rollout.  -> (ego observation after it, episode starts), both tensors that keep their address."""
        penv, ego, partner = self.penv, self.ego, self.penv.partner
        partner.maybe_train()                         # the partner's buffer filled up during the previous rollout
        n = ego.buffer.n_steps
        if self.graph is None and self.replays >= 1:  # the first rollout ran eagerly (warm-up); capture the second
            torch.cuda.synchronize(penv.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._body()
            self.graph = g
            # capture records, it does not execute: rewind the Python-side counters the body advanced
            partner.num_timesteps -= n * penv.num_envs
        if self.graph is not None:
            self.graph.replay()
            ego.buffer.pos = n
            partner.model.buffer.pos = n
            partner.num_timesteps += n * penv.num_envs
        else:
            self._body()
        self.replays += 1
        return penv._ring[0][:, 0], self.starts


class SB3VecEnvAdapter:
    """Stable-Baselines3 `VecEnv` duck type over `PantheonVecEnv` (interface 2):
    numpy in / numpy out, `step_async` / `step_wait`, auto-reset with
    `infos[i]["terminal_observation"]`, `get_attr` / `set_attr` / `env_method` / `env_is_wrapped` /
    `seed` / `close`.  This is what replaces the `DummyVecEnv` of ONE env that
    `collect_rollouts` steps (sb3_contrib/ppo_recurrent/ppo_recurrent.py:233-252) for a learner that
    keeps its rollout buffers on the host; a learner living on the GPU should use
    `PantheonVecEnv` / `OvercookedVecEnv` directly and never leave the device.
    SB3 itself is not imported (it is not installed in this image)."""

    def __init__(self, penv: PantheonVecEnv, dict_obs: bool = False):
        self.penv = penv
        self.num_envs = penv.num_envs
        self.observation_space = penv.env.observation_space
        self.action_space = penv.env.action_space
        self.dict_obs = dict_obs
        self._actions = None
        self.render_mode = None

    def _obs(self, t: torch.Tensor):
        a = t.cpu().numpy()
        if not self.dict_obs:
            return a
        return {k: a[:, s] for k, s in self.penv.env.obs_layout.items()}

    def reset(self):
        return self._obs(self.penv.reset())

    def step_async(self, actions):
        self._actions = torch.as_tensor(actions, dtype=torch.int32, device=self.penv.device).reshape(self.num_envs, 2)

    def step_wait(self):
        obs, rew, done = self.penv.step(self._actions)
        d = done.bool().cpu().numpy()
        infos = [{} for _ in range(self.num_envs)]
        if d.any():
            term = self.penv.terminal_obs.cpu().numpy()
            for i in d.nonzero()[0]:
                infos[i]["terminal_observation"] = (term[i] if not self.dict_obs else
                                                    {k: term[i, s] for k, s in self.penv.env.obs_layout.items()})
                infos[i]["TimeLimit.truncated"] = False
        return self._obs(obs), rew.cpu().numpy().astype("float32"), d, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.penv.env.close()

    def seed(self, seed=None):
        return [None] * self.num_envs          # placements / actions are seeded at construction (oc_config.seed)

    def get_attr(self, attr_name, indices=None):
        n = self.num_envs if indices is None else len(list(indices) if not isinstance(indices, int) else [indices])
        return [getattr(self.penv.env, attr_name)] * n

    def set_attr(self, attr_name, value, indices=None):
        raise AttributeError("the batched env shares one configuration; rebuild it to change %r" % attr_name)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        if method_name == "render":
            idx = range(self.num_envs) if indices is None else ([indices] if isinstance(indices, int) else indices)
            return [self.penv.env.render(i) for i in idx]
        raise AttributeError("no per-env method %r" % method_name)

    def env_is_wrapped(self, wrapper_class, indices=None):
        n = self.num_envs if indices is None else len(list(indices) if not isinstance(indices, int) else [indices])
        return [False] * n
