"""TEST / MEASUREMENT INFRASTRUCTURE -- times the UNMODIFIED reference (kyle-he/gym-comm) on this
container's host cores, the way BASELINE.md section 3 describes: `OvercookedMultiEnv.multi_step`
incl. both observations, one env per process, one worker per core, uniform random (nav, comm)
actions, auto-reset inside the clock, stdout discarded, PYTHONHASHSEED=0.

    PYTHONHASHSEED=0 python -m oracle.time_reference [seconds] > profiles/r1_reference_cpu_build_container.json

Only runs where /root/reference exists (the build container); the GPU box times the oracle port instead."""
import json
import multiprocessing as mp
import os
import random
import sys
import time

CONFIGS = {
    "cfg1/cfg2 open-divider_tomato C=10 T=500": dict(level="open-divider_tomato", max_num_timesteps=500, num_communication=10),
    "cfg3 level partial-divider_salad (2 agents via the wrapper) T=500": dict(level="partial-divider_salad", max_num_timesteps=500),
    "cfg4 env_args20on_allergic": dict(level="random-open-divider_salad_small_cramped", max_num_timesteps=900,
                                       num_communication=8, fow_radius=10,
                                       ego_config=dict(CAN_MOVE=False, ALLERGIC=True, BLIND=False),
                                       partner_config=dict(CAN_MOVE=True, ALLERGIC=False, BLIND=True)),
    "cfg5 random-salad-superwide C=100 T=900": dict(level="random-salad-superwide", max_num_timesteps=900, num_communication=100),
}


def worker(args):
    name, seconds, seed = args
    from oracle import ref_harness
    ns = ref_harness.make_namespace(**CONFIGS[name])
    ref = ref_harness.LiveReference(ns, py_random_seed=seed)
    rng = random.Random(seed)
    C = ns.num_communication
    for _ in range(50):
        ref.step([rng.randrange(4), rng.randrange(4)], [rng.randrange(C), rng.randrange(C)])
    steps, resets = 0, 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        _, done = ref.step([rng.randrange(4), rng.randrange(4)], [rng.randrange(C), rng.randrange(C)])
        steps += 1
        if done:
            ref.reset()
            resets += 1
    return steps, resets, time.perf_counter() - t0


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
    cores = os.cpu_count() or 1
    out = {"what": "unmodified reference, OvercookedMultiEnv.multi_step incl. 2 observations and resets",
           "host": "build container", "cores": cores, "seconds_per_worker": seconds, "results": {}}
    for name in CONFIGS:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(worker, [(name, seconds, 100 + i) for i in range(cores)])
        steps = sum(r[0] for r in res)
        dt = max(r[2] for r in res)
        out["results"][name] = {"env_steps_per_s_all_cores": steps / dt, "agent_steps_per_s_all_cores": 2 * steps / dt,
                                "agent_steps_per_s_per_core": 2 * steps / dt / cores, "resets": sum(r[1] for r in res)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
