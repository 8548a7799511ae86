"""One-off long-horizon soak: fused device rollout vs the C oracle twin (same Philox actions, same
device-RNG resets) for tens of thousands of steps; rewards / dones every step, observations and packed
state summaries at chunk boundaries.  python tools/soak_parity.py [steps] [obs_every]

`obs_every` (default 10): every that many chunks ALL observations of the chunk's 50 steps are compared too -- inside a
launch the fused kernel never clears its single-pass float rows, so a stale feature would show up here."""
import argparse
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from gym_comm_b200 import levels_data  # noqa: E402
from gym_comm_b200.vec_env import OvercookedVecEnv  # noqa: E402
from oracle.c_oracle import COracle  # noqa: E402

D = dict(CAN_MOVE=True, ALLERGIC=False, BLIND=False)
CASES = {
    "cfg2": dict(level="open-divider_tomato", num_agents=2, max_num_timesteps=500, communication_on=True,
                 num_communication=10, ego_led=False, fow_radius=2, ego_config=D, partner_config=D),
    "cfg4": dict(level="random-open-divider_salad_small_cramped", num_agents=2, max_num_timesteps=900,
                 communication_on=True, num_communication=8, ego_led=False, fow_radius=10,
                 ego_config=dict(CAN_MOVE=False, ALLERGIC=True, BLIND=False),
                 partner_config=dict(CAN_MOVE=True, ALLERGIC=False, BLIND=True)),
    "salad3": dict(level="partial-divider_salad", num_agents=3, max_num_timesteps=400, communication_on=True,
                   num_communication=10, ego_led=False, fow_radius=2, ego_config=D, partner_config=D),
}
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
obs_every = int(sys.argv[2]) if len(sys.argv) > 2 else 10
E, chunk = 16384, 50
for name, cfg in CASES.items():
    text = levels_data.LEVELS[cfg["level"]]
    sub = levels_data.SUBTASKS[tuple(text.split("\n\n")[1].split("\n"))]
    env = OvercookedVecEnv(argparse.Namespace(**cfg), num_envs=E, device="cuda:0", seed=99)
    ora = COracle(text, sub, E, seed=99, **{k: v for k, v in cfg.items() if k != "level"})
    n, F = cfg["num_agents"], env.obs_width
    obs = torch.zeros((chunk, E, n, F), device="cuda:0")
    rew = torch.zeros((chunk, E, n), device="cuda:0")
    done = torch.zeros((chunk, E), dtype=torch.uint8, device="cuda:0")
    t0 = time.time()
    events = 0
    for c in range(steps // chunk):
        env.rollout(chunk, obs_out=obs, rew_out=rew, done_out=done)
        last = (c % 40 == 0) or c == steps // chunk - 1
        check_obs = c % obs_every == 0
        oo, orr, od, _ = ora.rollout(chunk, want_obs=check_obs)
        assert torch.equal(done.cpu(), torch.from_numpy(od)), (name, c)
        if check_obs:
            assert torch.equal(obs.cpu(), torch.from_numpy(oo.astype(np.float32))), (name, c, "observations")
        assert torch.equal(rew.cpu()[:, :, 0], torch.from_numpy(orr.astype(np.float32))), (name, c)
        events += int((orr.astype(np.float32) > -1.0).sum())
        if last:
            st, os_ = env.decode_state(), ora.state()
            assert np.array_equal(st["t"], os_["t"]) and np.array_equal(st["episodes"], os_["episodes"]), (name, c)
            assert np.array_equal(np.stack([st["agent_x"], st["agent_y"]], -1), os_["agents"]), (name, c)
    print("%s: %d steps x %d envs identical (%.0f s), episodes/env %.1f, high-reward steps %d" %
          (name, steps, E, time.time() - t0, float(env.decode_state()["episodes"].mean()), events), flush=True)
    env.close()
    ora.close()
print("soak ok")
