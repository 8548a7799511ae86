"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- the CPU arm timed beside the GPU numbers.

The reference itself (pure Python, /root/reference) does not exist on the GPU box, so the CPU
arm is the oracle PORT: ``kind="c"`` = oracle/oc_oracle.c, one env per thread-chunk over all host
cores (the strongest CPU statement of the path we have); ``kind="py"`` = oracle/spec_model.py,
one env per process -- structurally what the reference does (per-object Python loop), minus its
template-object and networkx overheads.  BASELINE.md records the reference's own measured
speed in the build container (1.2-1.7 k env-steps/s/core).
Same action distribution as the GPU run: nav ~ U{0..3}, comm ~ U{0..C-1}, auto-reset, resets
inside the clock.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import random
import time


def _level_text(ns):
    from gym_comm_b200 import levels_data   # data only (level layouts + canonical subtask order)
    text = levels_data.LEVELS[ns.level]
    recipes = tuple(text.split("\n\n")[1].split("\n"))
    return text, levels_data.SUBTASKS[recipes]


def _py_worker(args):
    ns_dict, seconds, seed = args
    import argparse
    from oracle.spec_model import SpecEnv
    ns = argparse.Namespace(**ns_dict)
    text, subtasks = _level_text(ns)
    rng = random.Random(seed)
    probe = SpecEnv.__new__(SpecEnv)
    probe.n = ns.num_agents
    probe._parse_level(text)

    def pl():
        return rng.sample(probe.counters, len(probe.random_reps)) if probe.random_reps else None
    env = SpecEnv(text, subtasks, num_agents=ns.num_agents, max_num_timesteps=ns.max_num_timesteps,
                  communication_on=ns.communication_on, num_communication=ns.num_communication,
                  ego_led=ns.ego_led, fow_radius=ns.fow_radius, ego_config=ns.ego_config,
                  partner_config=ns.partner_config, placements=pl())
    n, C = ns.num_agents, ns.num_communication
    steps = 0
    t_end = time.perf_counter() + seconds
    t0 = time.perf_counter()
    while time.perf_counter() < t_end:
        for _ in range(50):
            navs = [rng.randrange(4) for _ in range(n)]
            comms = [rng.randrange(C) for _ in range(n)]
            _, done, _ = env.step(navs, comms)
            for k in range(n):
                env.flat_obs(k)
            if done:
                env.reset(pl())
                for k in range(n):
                    env.flat_obs(k)
            steps += 1
    return steps, time.perf_counter() - t0


def run_all_cores(workload, ns, seconds, kind="py"):
    """Primary figure: the Python port (structurally the reference: a per-object Python loop, one
    env per process, all host cores).  The C port on all cores is reported next to it as
    ``c_port`` -- a far stronger CPU statement of the path than the reference itself."""
    cores = os.cpu_count() or 1
    n = ns.num_agents
    ns_dict = dict(vars(ns))
    py_seconds = seconds * 0.75 if kind == "py" else seconds * 0.25
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_py_worker, [(ns_dict, py_seconds, 1000 + i) for i in range(cores)])
    steps = sum(r[0] for r in res)
    dt = max(r[1] for r in res)
    py = {"value": steps * n / dt, "unit": "agent-steps/s", "cores": cores, "kind": "port",
          "impl": "oracle/spec_model.py (Python restatement of the reference path, one env per process)",
          "sample": "%.1f s per worker of random-action stepping incl. %d observations per step and resets" % (dt, n),
          "env_steps": steps, "seconds": dt}
    try:
        from oracle import c_oracle
        csteps, cdt, threads = c_oracle.throughput(ns, max(1.0, seconds - py_seconds))
        cport = {"value": csteps * n / cdt, "unit": "agent-steps/s", "cores": threads, "kind": "port",
                 "impl": "oracle/oc_oracle.c (C restatement, one thread per host core, %d envs per thread)" % c_oracle.ENVS_PER_THREAD,
                 "sample": "%.1f s of random-action stepping incl. obs featurisation and resets" % cdt,
                 "env_steps": csteps, "seconds": cdt}
    except Exception as ex:
        cport = {"error": repr(ex)}
    out = dict(py if kind == "py" else cport)
    out["c_port" if kind == "py" else "python_port"] = cport if kind == "py" else py
    out["reference_measured_in_build_container"] = (
        "the unmodified reference: 1.2-1.7 k env-steps/s per core (BASELINE.md section 2); it is pure Python and "
        "cannot travel to the GPU box")
    return out
