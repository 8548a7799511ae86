"""A/B sweep of the per-step kernels' host-side knobs in ONE process (GPU box): every variant = a set of OC_* environment
variables read by oc_create.  Times a CUDA graph of 32 back-to-back steps (float rows: oc_step; compact rows: oc_step_i8).

    python tools/step_sweep.py [workload ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, workload_namespace  # noqa: E402
from gym_comm_b200.vec_env import OvercookedVecEnv  # noqa: E402

VARIANTS = [{}, {"OC_PDL": "0"}, {"OC_ROW_BUFS": "2"}, {"OC_BLOCK_THREADS": "128"}, {"OC_BLOCK_THREADS": "96"},
            {"OC_BLOCK_THREADS": "64"}, {"OC_BLOCK_THREADS": "256"}, {"OC_ROW_ENVS": "16"}, {"OC_TMA": "0"}]
VARIANTS8 = [{}, {"OC_PDL": "0"}, {"OC_BLOCK_THREADS_I8": "128"}, {"OC_BLOCK_THREADS_I8": "64"}, {"OC_BLOCK_THREADS_I8": "256"},
             {"OC_TMA": "0"}]
KEYS = sorted({k for v in VARIANTS + VARIANTS8 for k in v})


def timed(fn_step, R=32, reps=40):
    for i in range(R):
        fn_step(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(R):
            fn_step(i)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(int(2e6))
        a.record()
        for _ in range(reps):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e3 / (reps * R))
    return best


def main():
    names = sys.argv[1:] or ["cfg2"]
    dev = torch.device("cuda", 0)
    for name in names:
        w = WORKLOADS[name]
        E = w["envs"]
        ns = workload_namespace(w)
        A = ns.num_agents
        R = 32
        acts = torch.stack([torch.randint(0, 4, (R, E, A), device=dev, dtype=torch.int32),
                            torch.randint(0, w["num_communication"], (R, E, A), device=dev, dtype=torch.int32)], -1).contiguous()
        acts8 = acts.to(torch.uint8)
        for compact, variants in ((False, VARIANTS), (True, VARIANTS8)):
            for v in variants:
                for k in KEYS:
                    os.environ.pop(k, None)
                os.environ.update(v)
                try:
                    env = OvercookedVecEnv(ns, num_envs=E, device=dev, seed=1)
                    F = env.obs_width
                    env.reset()
                    if not compact:
                        nslot = max(2, min(R, int(1.5e9 // (E * A * F * 4))))
                        obs = torch.empty((nslot, E, A, F), device=dev)
                        rew = torch.empty((R, E, A), device=dev)
                        done = torch.empty((R, E), dtype=torch.uint8, device=dev)
                        t = timed(lambda i: env.step(acts[i % R], obs_out=obs[i % nslot], rew_out=rew[i % R], done_out=done[i % R]), R)
                        del obs
                    else:
                        o8 = torch.empty((R, E, A, F - 1), dtype=torch.int8, device=dev)
                        ts = torch.empty((R, E), device=dev)
                        rew = torch.empty((R, E), device=dev)
                        done = torch.empty((R, E), dtype=torch.uint8, device=dev)
                        t = timed(lambda i: env.step_i8(acts8[i % R], o8[i % R], ts[i % R], rew_out=rew[i % R], done_out=done[i % R]), R)
                        del o8
                    print("%s %-8s %-32s %7.2f us/step  %6.2f G agent-steps/s" % (name, "compact" if compact else "float", v or "default", t, E * A / t / 1e3), flush=True)
                    env.close()
                    torch.cuda.empty_cache()
                except Exception as ex:
                    print(name, "compact" if compact else "float", v, "FAILED", repr(ex)[:200], flush=True)


if __name__ == "__main__":
    main()
