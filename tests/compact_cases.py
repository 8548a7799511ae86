"""Cases for the compact-row kernels (`oc_step_i8` / `oc_reset_i8`: int8 rows + f32 clock written by the step /
reset kernel itself), shared between the GPU suite (CUDA library) and the CPU suite (emulation of the same device
code).  The float rows of a twin env stepped with the same actions are the reference: they are themselves checked
against the oracles and the golden traces elsewhere, and `obs_i8.astype(f32)` with the clock appended must BE
the float row."""
import argparse

import numpy as np
import torch

D = dict(CAN_MOVE=True, ALLERGIC=False, BLIND=False)
CASES = [("open-divider_tomato", 2, 13, 10, 2), ("random-salad-superwide", 2, 11, 100, 2),
         ("partial-divider_salad", 3, 9, 6, 1), ("random-open-divider_salad_small_cramped", 2, 12, 8, 10),
         ("open-divider_salad", 4, 10, 7, 3)]


def run_step_i8_equals_float_rows(make_env, device, level, A, T, C, fow, E, u8_actions, per_env_reward, steps=None):
    ns = argparse.Namespace(level=level, num_agents=A, max_num_timesteps=T, communication_on=True, num_communication=C,
                            ego_led=False, fow_radius=fow, ego_config=D,
                            partner_config=dict(D, BLIND=(level.startswith("random-open"))))
    a_env = make_env(ns, num_envs=E, seed=17, auto_reset=True)
    b_env = make_env(ns, num_envs=E, seed=17, auto_reset=True)
    F = a_env.obs_width
    rng = np.random.default_rng(E)
    o8, ts = b_env.compact_buffers()
    f = a_env.reset()
    b_env.reset_i8(o8, ts)
    assert torch.equal(o8.to(torch.float32), f[..., :-1]) and torch.equal(ts, f[:, 0, -1])
    term_f = torch.zeros((E, A, F), device=device)
    term8 = torch.zeros((E, A, F - 1), dtype=torch.int8, device=device)
    term_ts = torch.zeros((E,), device=device)
    rew_b = torch.zeros((E,) if per_env_reward else (E, A), device=device)
    done_b = torch.zeros((E,), dtype=torch.uint8, device=device)
    nfin = 0
    for t in range(steps or (2 * T + 5)):
        a = np.stack([rng.integers(0, 4, (E, A)), rng.integers(0, C, (E, A))], -1)
        a32 = torch.from_numpy(a.astype(np.int32)).to(device)
        ab = torch.from_numpy(a.astype(np.uint8)).to(device) if u8_actions else a32
        f, rew, done = a_env.step(a32, term_obs_out=term_f, want_f64=True)
        b_env.step_i8(ab, o8, ts, rew_out=rew_b, done_out=done_b, term_obs_out=term8, term_timestep_out=term_ts, want_f64=True)
        assert torch.equal(done_b, done), t
        assert torch.equal(rew_b if per_env_reward else rew_b[:, 0], rew[:, 0]), t
        assert torch.equal(b_env.rewards64, a_env.rewards64), t
        assert torch.equal(o8.to(torch.float32), f[..., :-1]), t
        assert torch.equal(ts, f[:, 0, -1]), t
        d = done.bool()
        nfin += int(d.sum())
        assert torch.equal(term8.to(torch.float32)[d], term_f[..., :-1][d]) and torch.equal(term_ts[d], term_f[:, 0, -1][d]), t
        if t == 3:          # masked reset in both formats
            m = torch.from_numpy((np.arange(E) % 3 == 0).astype(np.uint8)).to(device)
            f = a_env.reset(mask=m)
            b_env.reset_i8(o8, ts, mask=m)
            assert torch.equal(o8.to(torch.float32), f[..., :-1]) and torch.equal(ts, f[:, 0, -1])
    assert torch.equal(a_env.get_state(), b_env.get_state())
    assert nfin >= E
    # rows of envs that never finished stay untouched; finished ones hold their last terminal row in both formats
    assert torch.equal(term8.to(torch.float32), term_f[..., :-1])
    a_env.close()
    b_env.close()


def run_set_state_sanitises(make_env, device):
    """oc_set_state with words from nowhere: agent / object cells beyond the grid are clamped, holders beyond the
    agents released, empty slots normalised -- and the env steps on without touching memory it does not own."""
    ns = argparse.Namespace(level="open-divider_tomato", num_agents=2, max_num_timesteps=50, communication_on=True,
                            num_communication=10, ego_led=False, fow_radius=2, ego_config=D, partner_config=D)
    E = 70
    env = make_env(ns, num_envs=E, seed=1, auto_reset=True)
    env.reset()
    g = torch.Generator().manual_seed(5)
    junk = torch.randint(-2 ** 31, 2 ** 31 - 1, (E, 16), generator=g, dtype=torch.int64).to(torch.int32).to(device)
    junk[:, 0] = junk[:, 0] & 0x1F                      # keep the clock small so the episode goes on
    env.set_state(junk)
    dec = env.decode_state()
    ncell = env.level.width * env.level.height
    assert (dec["agent_cell"] < ncell).all()
    alive = dec["obj_contents"] != 0
    assert (dec["obj_cell"][alive] < ncell).all()
    assert np.isin(dec["obj_holder"][alive], [0, 1, 7]).all()
    st = env.get_state().cpu().numpy().view(np.uint32)
    assert (st[:, 8:14][~alive] == 0x00FF0700).all()
    a = torch.zeros((E, 2, 2), dtype=torch.int32, device=device)
    for t in range(60):
        a[..., 0] = t % 4
        env.step(a)
    assert torch.isfinite(env.obs).all()
    # a state exported by the env itself survives the round trip unchanged
    st0 = env.get_state()
    env.set_state(st0)
    assert torch.equal(env.get_state(), st0)
    env.close()
