"""GPU: chained steps (OC_FLAG_CHAIN_HEAD / OC_FLAG_CHAINED) -- overlapping launches that depend on each other through
the chain counter instead of on grid completion -- give exactly the results of plain stepping: eagerly, inside a
replayed CUDA graph, with several chunks per thread block, in the compact format; and every misuse fails loudly."""
import argparse

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
D = dict(CAN_MOVE=True, ALLERGIC=False, BLIND=False)


def _ns(level, A, T, C):
    return argparse.Namespace(level=level, num_agents=A, max_num_timesteps=T, communication_on=True, num_communication=C,
                              ego_led=False, fow_radius=2, ego_config=D, partner_config=D)


def _actions(n, E, A, C, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.stack([torch.randint(0, 4, (n, E, A), generator=g, device=DEV, dtype=torch.int32),
                        torch.randint(0, C, (n, E, A), generator=g, device=DEV, dtype=torch.int32)], -1).contiguous()


@pytest.mark.parametrize("level,A,T,C,E", [("open-divider_tomato", 2, 23, 10, 65536), ("partial-divider_salad", 3, 17, 10, 300000),
                                           ("random-salad-superwide", 2, 19, 100, 70001), ("open-divider_tomato", 2, 9, 4, 33),
                                           # rows of 96 floats: grouped rows, one bulk copy per group of four
                                           ("random-open-divider_salad_small_cramped", 2, 21, 8, 65569)])
def test_chained_steps_equal_plain_steps(level, A, T, C, E):
    from gym_comm_b200.vec_env import OvercookedVecEnv
    n = 36
    plain = OvercookedVecEnv(_ns(level, A, T, C), num_envs=E, device=DEV, seed=5, auto_reset=True)
    chain = OvercookedVecEnv(_ns(level, A, T, C), num_envs=E, device=DEV, seed=5, auto_reset=True)
    F = plain.obs_width
    acts = _actions(n, E, A, C, 1)
    R = 3                                                  # output slots: consecutive steps never share one
    obs = torch.zeros((R, E, A, F), device=DEV)
    rew = torch.zeros((R, E, A), device=DEV)
    done = torch.zeros((R, E), dtype=torch.uint8, device=DEV)
    want_o, want_r, want_d = [], [], []
    plain.reset()
    chain.reset()
    for i in range(n):
        o, r, d = plain.step(acts[i])
        want_o.append(o.clone()); want_r.append(r.clone()); want_d.append(d.clone())

    def run_chain(check_from):
        for i in range(n):
            chain.step(acts[i], obs_out=obs[i % R], rew_out=rew[i % R], done_out=done[i % R], chain="head" if i == 0 else "next")
            if i >= check_from and (i % R == R - 1 or i == n - 1):       # read the last R slots back (a host sync in the
                torch.cuda.synchronize()                                  # middle of a chain is harmless)
                for j in range(max(0, i - R + 1), i + 1):
                    assert torch.equal(obs[j % R], want_o[j]) and torch.equal(rew[j % R], want_r[j]) and torch.equal(done[j % R], want_d[j]), j
    # eager chain: the whole run is one chain; only the tail is compared (slots are overwritten as the chain advances)
    run_chain(check_from=n - R)
    assert torch.equal(chain.get_state(), plain.get_state())
    # the same run inside a CUDA graph, replayed: head first, everything else chained
    chain.reset(); plain.reset()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            chain.step(acts[i], obs_out=obs[i % R], rew_out=rew[i % R], done_out=done[i % R], chain="head" if i == 0 else "next")
    for rep in range(3):
        g.replay()
        for i in range(n):
            o, r, d = plain.step(acts[i])
        torch.cuda.synchronize()
        assert torch.equal(obs[(n - 1) % R], o) and torch.equal(rew[(n - 1) % R], r) and torch.equal(done[(n - 1) % R], d), rep
        assert torch.equal(chain.get_state(), plain.get_state()), rep
    plain.close(); chain.close()


def test_chained_compact_steps_equal_plain_steps():
    from gym_comm_b200.vec_env import OvercookedVecEnv
    E, A, T, C, n = 50001, 2, 21, 10, 30
    plain = OvercookedVecEnv(_ns("open-divider_tomato", A, T, C), num_envs=E, device=DEV, seed=9, auto_reset=True)
    chain = OvercookedVecEnv(_ns("open-divider_tomato", A, T, C), num_envs=E, device=DEV, seed=9, auto_reset=True)
    F = plain.obs_width
    acts = _actions(n, E, A, C, 2).to(torch.uint8)
    o8 = [torch.zeros((E, A, F - 1), dtype=torch.int8, device=DEV) for _ in range(2)]     # own allocations: 16-byte aligned
    ts = [torch.zeros((E,), device=DEV) for _ in range(2)]
    rew = [torch.zeros((E,), device=DEV) for _ in range(2)]
    done = [torch.zeros((E,), dtype=torch.uint8, device=DEV) for _ in range(2)]
    po, pt = plain.compact_buffers()
    pr, pd = torch.zeros(E, device=DEV), torch.zeros(E, dtype=torch.uint8, device=DEV)
    for i in range(n):
        plain.step_i8(acts[i], po, pt, rew_out=pr, done_out=pd)
        chain.step_i8(acts[i], o8[i % 2], ts[i % 2], rew_out=rew[i % 2], done_out=done[i % 2], chain="head" if i == 0 else "next")
    torch.cuda.synchronize()
    k = (n - 1) % 2
    assert torch.equal(o8[k], po) and torch.equal(ts[k], pt) and torch.equal(rew[k], pr) and torch.equal(done[k], pd)
    assert torch.equal(chain.get_state(), plain.get_state())
    plain.close(); chain.close()


def test_chain_misuse_fails_loudly():
    from gym_comm_b200.vec_env import OvercookedVecEnv
    E = 1000
    env = OvercookedVecEnv(_ns("open-divider_tomato", 2, 30, 10), num_envs=E, device=DEV, seed=1, auto_reset=True)
    a = _actions(4, E, 2, 10, 3)
    other = torch.zeros_like(env.obs), torch.zeros_like(env.rewards), torch.zeros_like(env.dones)
    env.reset()
    with pytest.raises(RuntimeError, match="OC_FLAG_CHAINED"):          # no head
        env.step(a[0], chain="next")
    env.step(a[0], chain="head")
    with pytest.raises(RuntimeError, match="different"):                # same output buffers as the previous step
        env.step(a[1], chain="next")
    env.step(a[1], obs_out=other[0], rew_out=other[1], done_out=other[2], chain="next")
    env.step(a[2], chain="next")                                        # alternating buffers: fine
    env.reset()
    with pytest.raises(RuntimeError, match="OC_FLAG_CHAINED"):          # a reset ended the chain
        env.step(a[3], obs_out=other[0], rew_out=other[1], done_out=other[2], chain="next")
    env.step(a[3])
    with pytest.raises(RuntimeError, match="OC_FLAG_CHAINED"):          # so did a plain step
        env.step(a[3], obs_out=other[0], rew_out=other[1], done_out=other[2], chain="next")
    torch.cuda.synchronize()
    env.close()
