"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- the CPU arm timed beside the GPU numbers.

Three CPU statements of the path, all with the GPU run's action distribution (nav ~ U{0..3},
comm ~ U{0..C-1}), auto-reset, resets inside the clock, one worker per host core:

``reference``  the UNMODIFIED reference (`OvercookedMultiEnv.multi_step`, gym_comm/envs/overcooked_env.py:207-282),
               one env per process, PYTHONHASHSEED=0, stdout discarded -- oracle/time_reference.py run as a
               subprocess.  Found at /root/reference (build container) or in the git-ignored copy
               oracle/stage_ref.py staged under oracle/_ref/ (GPU box).  This is the primary figure
               (`kind: "reference"`) whenever it is available.
``python_port``  oracle/spec_model.py, one env per process -- structurally the reference's per-object Python
               loop minus its template-object and networkx overheads (`kind: "port"`).
``c_port``     oracle/oc_oracle.c, one thread per core -- a far stronger CPU statement than the reference.
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


def run_reference(workload, seconds):
    """The unmodified reference on all cores, or None when no reference tree is reachable."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = {k: v for k, v in workload.items() if k != "envs"}
    env = dict(os.environ, PYTHONHASHSEED="0")
    try:
        out = subprocess.run([sys.executable, "-m", "oracle.time_reference", "--seconds", "%g" % seconds,
                              "--workload", json.dumps(cfg)], cwd=root, env=env, stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, timeout=seconds * 6 + 120, check=True).stdout.decode()
        d = json.loads(out.strip().splitlines()[-1])
    except Exception as ex:
        return {"error": repr(ex)}
    if not d.get("available"):
        return None
    n = int(workload.get("num_agents", 2))
    return {"value": d["agent_steps_per_s"], "unit": "agent-steps/s", "cores": d["cores"], "kind": "reference",
            "impl": "the unmodified reference (OvercookedMultiEnv.multi_step incl. observations), one env per process, "
                    "PYTHONHASHSEED=0, from %s" % d["reference_root"],
            "sample": "%.1f s per worker of random-action stepping incl. %d observations per step and %d resets" %
                      (d["seconds"], n, d["resets"]),
            "env_steps": d["env_steps"], "seconds": d["seconds"], "per_core": d["agent_steps_per_s_per_core"]}


def run_python_port(ns, seconds):
    cores = os.cpu_count() or 1
    n = ns.num_agents
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_py_worker, [(dict(vars(ns)), seconds, 1000 + i) for i in range(cores)])
    steps = sum(r[0] for r in res)
    dt = max(r[1] for r in res)
    return {"value": steps * n / dt, "unit": "agent-steps/s", "cores": cores, "kind": "port",
            "impl": "oracle/spec_model.py (Python restatement of the reference path, one env per process)",
            "sample": "%.1f s per worker of random-action stepping incl. %d observations per step and resets" % (dt, n),
            "env_steps": steps, "seconds": dt}


def run_c_port(ns, seconds):
    try:
        from oracle import c_oracle
        n = ns.num_agents
        csteps, cdt, threads = c_oracle.throughput(ns, max(1.0, seconds))
        return {"value": csteps * n / cdt, "unit": "agent-steps/s", "cores": threads, "kind": "port",
                "impl": "oracle/oc_oracle.c (C restatement, one thread per host core, %d envs per thread)" % c_oracle.ENVS_PER_THREAD,
                "sample": "%.1f s of random-action stepping incl. obs featurisation and resets" % cdt,
                "env_steps": csteps, "seconds": cdt}
    except Exception as ex:
        return {"error": repr(ex)}


def run_all_cores(workload, ns, seconds):
    """Primary figure: the unmodified reference when a reference tree is reachable (`kind: "reference"`),
    else the Python port (`kind: "port"`); the other statements ride along under their own keys."""
    ref = run_reference(workload, seconds * 0.6)
    have_ref = isinstance(ref, dict) and "value" in ref
    py = run_python_port(ns, seconds * (0.25 if have_ref else 0.75))
    cport = run_c_port(ns, seconds * (0.15 if have_ref else 0.25))
    out = dict(ref if have_ref else py)
    if have_ref:
        out["python_port"] = py
    elif ref is not None:
        out["reference_error"] = ref.get("error")
    out["c_port"] = cport
    return out
