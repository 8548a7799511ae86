"""Can the step kernel write its outputs straight into page-locked HOST memory, and how fast?  (GPU box)

    python tools/zero_copy_probe.py [workload]

Under unified virtual addressing a cudaHostAlloc'd pointer is valid on the device, so `oc_step` can be handed a
pinned host buffer as `obs` / `rew` / `done`: the copy engine's bulk stores then travel over PCIe while the
kernel runs and no cudaMemcpyAsync follows.  Prints, per step: device buffers + one D2H copy of the float rows
vs the kernel writing to host memory directly, plus the raw pinned-copy bandwidth of this box."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, workload_namespace  # noqa: E402
from gym_comm_b200.vec_env import OvercookedVecEnv  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    w = WORKLOADS[name]
    E = w["envs"]
    dev = torch.device("cuda", 0)
    env = OvercookedVecEnv(workload_namespace(w), num_envs=E, device=dev, seed=1)
    A, F = env.num_agents, env.obs_width
    acts = torch.stack([torch.randint(0, 4, (8, E, A), device=dev, dtype=torch.int32),
                        torch.randint(0, w["num_communication"], (8, E, A), device=dev, dtype=torch.int32)], -1).contiguous()
    obs_d = torch.empty((E, A, F), device=dev)
    rew_d = torch.empty((E, A), device=dev)
    done_d = torch.empty((E,), dtype=torch.uint8, device=dev)
    obs_h = torch.empty((E, A, F), pin_memory=True)
    rew_h = torch.empty((E, A), pin_memory=True)
    done_h = torch.empty((E,), dtype=torch.uint8, pin_memory=True)
    lib, h = env.lib, env._handle
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    env.reset()
    N = 50

    # raw pinned copy bandwidth
    big = torch.empty(64 * 1024 * 1024, dtype=torch.uint8, device=dev)
    big_h = torch.empty(64 * 1024 * 1024, dtype=torch.uint8, pin_memory=True)
    for label, a, b_ in (("D2H", big_h, big), ("H2D", big, big_h)):
        a.copy_(b_, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            a.copy_(b_, non_blocking=True)
        torch.cuda.synchronize()
        print("pinned %s copy: %.1f GB/s (64 MiB x 10)" % (label, 10 * big.numel() / (time.perf_counter() - t0) / 1e9))

    def run(label, fn):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(N):
            fn(i)
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / N
        print("%-64s %8.1f us/step  %.3f G agent-steps/s" % (label, dt * 1e6, E * A / dt / 1e9))

    def staged(i):
        lib.check(lib.step(h, p(acts[i % 8]), p(obs_d), p(rew_d), None, p(done_d), None, 1, st), "oc_step")
        obs_h.copy_(obs_d, non_blocking=True)
        rew_h.copy_(rew_d, non_blocking=True)
        done_h.copy_(done_d, non_blocking=True)

    def direct(i):
        lib.check(lib.step(h, p(acts[i % 8]), p(obs_h), p(rew_h), None, p(done_h), None, 1, st), "oc_step")

    run("%s float rows: device buffers + 3 D2H copies (%.1f MB)" % (name, obs_h.numel() * 4 / 1e6), staged)
    try:
        run("%s float rows: kernel stores straight to pinned host memory" % name, direct)
        # same values?
        staged(0)
        torch.cuda.synchronize()
        ref = obs_h.clone()
        st0 = env.get_state()
        env.set_state(st0)
        print("direct path ran; obs checksum", float(obs_h.sum()), "ref", float(ref.sum()))
    except Exception as ex:
        print("direct path failed:", repr(ex))


if __name__ == "__main__":
    main()
