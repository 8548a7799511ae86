"""Where one launch of the step kernel spends its time (GPU box, probe build only):

    nvcc ... -DOC_PHASE_PROBE gym_comm_b200/csrc/oc_kernels.cu -o gym_comm_b200/variants/liboc_b200_probe.so
    OC_B200_LIB=$PWD/gym_comm_b200/variants/liboc_b200_probe.so python tools/probe_step.py [workload]

Lane 0 of every warp stamps %clock64 at the phase boundaries (see OC_PROBE in oc_kernels.cu) and
%globaltimer at entry / exit; the launch looked at is the LAST of a CUDA graph of back-to-back steps, i.e.
the steady state bench.py's `step_api` measures.  Prints the median / p90 length of each phase in SM cycles
and the chip-wide spread of the entry and exit times."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, workload_namespace  # noqa: E402
from gym_comm_b200 import _cabi  # noqa: E402
from gym_comm_b200.vec_env import OvercookedVecEnv  # noqa: E402

PHASES = ["entry->tables+clear issued", "griddepcontrol.wait (plain launches)", "tables landed (mbarrier)", "chain flag + state+actions arrive",
          "dynamics + reward/done", "finish/reset + state store", "obs clear/fill + bulk store issue",
          "wait for the copy engine to read the rows"]


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    chained = "--chain" in sys.argv
    name = args[0] if args else "cfg2"
    w = WORKLOADS[name]
    E, R = w["envs"], 16
    dev = torch.device("cuda", 0)
    lib = _cabi.default_library()
    setp = lib.lib.oc_debug_set_probe
    setp.restype, setp.argtypes = C.c_int, [C.c_void_p]
    env = OvercookedVecEnv(workload_namespace(w), num_envs=E, device=dev, seed=1)
    A, F = env.num_agents, env.obs_width
    obs = torch.empty((R, E, A, F), device=dev)
    rew = torch.empty((R, E, A), device=dev)
    done = torch.empty((R, E), dtype=torch.uint8, device=dev)
    acts = torch.stack([torch.randint(0, 4, (R, E, A), device=dev, dtype=torch.int32),
                        torch.randint(0, w["num_communication"], (R, E, A), device=dev, dtype=torch.int32)], -1).contiguous()
    probe = torch.zeros((4096 * 8 * 16,), dtype=torch.int64, device=dev)
    env.reset()
    for i in range(R):
        env.step(acts[i], obs_out=obs[i], rew_out=rew[i], done_out=done[i])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(R):
            env.step(acts[i], obs_out=obs[i], rew_out=rew[i], done_out=done[i],
                     chain=(("head" if i == 0 else "next") if chained else None))
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    assert setp(C.c_void_p(probe.data_ptr())) == 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    print("%s%s: %.2f us per step over 20 graph replays of %d steps (probe build)" %
          (name, " CHAINED" if chained else "", a.elapsed_time(b) * 1e3 / (20 * R), R))
    p = probe.cpu().numpy().reshape(-1, 16).astype(np.int64)
    p = p[p[:, 9] != 0]                        # warps that ran (every launch overwrites its own slots: last launch wins)
    print("warps recorded: %d on %d SMs" % (len(p), len(set(p[:, 11].tolist()))))
    c = p[:, :9]
    for k, label in enumerate(PHASES):
        d = c[:, k + 1] - c[:, k]
        print("  %-46s median %6d  p10 %6d  p90 %6d cycles" % (label, np.median(d), np.percentile(d, 10), np.percentile(d, 90)))
    tot = c[:, 8] - c[:, 0]
    print("  %-46s median %6d  p10 %6d  p90 %6d cycles" % ("whole warp", np.median(tot), np.percentile(tot, 10), np.percentile(tot, 90)))
    g0, g1 = p[:, 9], p[:, 10]
    base = g0.min()
    print("globaltimer (ns): first entry 0, last entry %d, first exit %d, last exit %d, median warp lifetime %d" %
          (g0.max() - base, g1.min() - base, g1.max() - base, np.median(g1 - g0)))
    # per-SM: when do its CTAs start relative to the first entry on the chip
    starts = sorted(set(((g0 - base) // 100 * 100).tolist()))
    print("distinct entry times (100 ns bins): %s" % starts[:24])


if __name__ == "__main__":
    main()
