"""TEST / MEASUREMENT INFRASTRUCTURE -- times the UNMODIFIED reference (kyle-he/gym-comm) on this host's
cores the way SURVEY section 8d / BASELINE.md section 3 describe: `OvercookedMultiEnv.multi_step` incl. both
observations (gym_comm/envs/overcooked_env.py:207-282), one env per process, one worker per core, uniform
random (nav, comm) actions, auto-reset inside the clock, stdout discarded, PYTHONHASHSEED=0.
(3-4 agents: the wrapper only drives two players, so `OvercookedEnvironment.step` + one
`get_observation2` per agent, exactly what the parity harness does.)

The reference is found by oracle/ref_harness.py: `/root/reference` in the build container, the unmodified
copy staged by oracle/stage_ref.py into the git-ignored `oracle/_ref/` on the GPU box.

    PYTHONHASHSEED=0 python -m oracle.time_reference --seconds 10 --workload '{"level": "open-divider_tomato", ...}'
    PYTHONHASHSEED=0 python -m oracle.time_reference --all --seconds 10     # the four BASELINE configs

Prints ONE JSON object.  `bench.py` runs it as a subprocess (the hash seed is fixed at interpreter start)."""
import argparse
import json
import multiprocessing as mp
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    "cfg2": dict(level="open-divider_tomato", num_agents=2, max_num_timesteps=500, num_communication=10, fow_radius=2),
    "cfg3": dict(level="partial-divider_salad", num_agents=3, max_num_timesteps=500, num_communication=10, fow_radius=2),
    "cfg4": dict(level="random-open-divider_salad_small_cramped", num_agents=2, max_num_timesteps=900,
                 num_communication=8, fow_radius=10,
                 ego_config=dict(CAN_MOVE=False, ALLERGIC=True, BLIND=False),
                 partner_config=dict(CAN_MOVE=True, ALLERGIC=False, BLIND=True)),
    "cfg5": dict(level="random-salad-superwide", num_agents=2, max_num_timesteps=900, num_communication=100, fow_radius=2),
}
NS_KEYS = ("level", "num_agents", "max_num_timesteps", "communication_on", "num_communication", "ego_led",
           "fow_radius", "ego_config", "partner_config")


def worker(args):
    cfg, seconds, seed = args
    from oracle import ref_harness
    ns = ref_harness.make_namespace(**{k: v for k, v in cfg.items() if k in NS_KEYS})
    ref = ref_harness.LiveReference(ns, py_random_seed=seed)
    rng = random.Random(seed)
    n, C = ns.num_agents, ns.num_communication

    def one_step():
        _, done = ref.step([rng.randrange(4) for _ in range(n)], [rng.randrange(C) for _ in range(n)])
        if n > 2:                      # the 2-player wrapper featurises inside multi_step; wider games do it here
            for k in range(n):
                ref.obs(k)
        return done
    for _ in range(20):
        one_step()
    steps, resets = 0, 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        if one_step():
            ref.reset()
            resets += 1
        steps += 1
    return steps, resets, time.perf_counter() - t0


def run(cfg, seconds, cores=None):
    cores = cores or os.cpu_count() or 1
    t0 = time.perf_counter()
    # import the reference (and what it drags in: torch, wandb) once, before forking the workers
    from oracle import ref_harness
    ref_harness.LiveReference(ref_harness.make_namespace(**{k: v for k, v in cfg.items() if k in NS_KEYS}))
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(worker, [(cfg, seconds, 100 + i) for i in range(cores)])
    steps = sum(r[0] for r in res)
    dt = max(r[2] for r in res)
    n = int(cfg.get("num_agents", 2))
    return {"env_steps": steps, "seconds": dt, "cores": cores, "resets": sum(r[1] for r in res),
            "agent_steps_per_s": steps * n / dt, "agent_steps_per_s_per_core": steps * n / dt / cores,
            "wall_s": time.perf_counter() - t0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--workload", default=None, help="JSON dict of Namespace fields")
    ap.add_argument("--all", action="store_true")
    ap.add_argument("--cores", type=int, default=0)
    a = ap.parse_args()
    from oracle import ref_harness
    out = {"what": "unmodified reference, OvercookedMultiEnv.multi_step incl. observations and resets, one env per process",
           "reference_root": ref_harness.REFERENCE_ROOT, "hashseed": os.environ.get("PYTHONHASHSEED"),
           "available": ref_harness.reference_available()}
    if not out["available"]:
        print(json.dumps(out))
        return 0
    if a.all:
        out["results"] = {k: run(v, a.seconds, a.cores or None) for k, v in CONFIGS.items()}
    else:
        cfg = json.loads(a.workload) if a.workload else CONFIGS["cfg2"]
        out.update(run(cfg, a.seconds, a.cores or None))
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
