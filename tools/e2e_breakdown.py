"""Where one host-buffer step (bench.py's `e2e`) spends its time, piece by piece (GPU box):

    python tools/e2e_breakdown.py [workload]

All wall-clock per step (perf_counter around call + synchronise, median of N), cfg2 unless told otherwise:
  A  compact step kernel alone, device buffers (oc_step_i8, u8 actions on the device)      -> launch + kernel + sync
  B  one pinned device->host copy of the block (6.49 MB at cfg2)                             -> the PCIe floor
  C  A + B on one stream (what oc_step_host_block enqueues, minus the action upload)
  D  oc_step_host_block through raw ctypes (pinned u8 actions read by the kernel)            -> the C ABI call
  E  OvercookedHostVecEnv(obs_format="i8").step()                                            -> + the Python class
  F  oc_step_i8 writing every output straight into pinned host memory (no copy at all)
  G  D with OC_FLAG_NO_SYNC, then oc_sync                                                     -> enqueue vs wait split
"""
import ctypes as C
import os
import statistics
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, workload_namespace  # noqa: E402
from gym_comm_b200 import _cabi  # noqa: E402
from gym_comm_b200.host_env import OvercookedHostVecEnv  # noqa: E402
from gym_comm_b200.vec_env import OvercookedVecEnv  # noqa: E402


def med(fn, n=200, warm=10):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    ts = []
    for i in range(n):
        t0 = time.perf_counter()
        fn(i)
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts) * 1e6, min(ts) * 1e6


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    w = WORKLOADS[name]
    E = w["envs"]
    ns = workload_namespace(w)
    dev = torch.device("cuda", 0)
    env = OvercookedVecEnv(ns, num_envs=E, device=dev, seed=1)
    A, F = env.num_agents, env.obs_width
    lib, h = env.lib, env._handle
    lay = _cabi.OcHostBlock()
    lib.check(lib.host_block_layout(h, C.byref(lay)), "layout")
    total = int(lay.total_bytes)
    blk_d = torch.zeros(total, dtype=torch.uint8, device=dev)
    blk_h = torch.zeros(total, dtype=torch.uint8, pin_memory=True)
    acts_d = torch.stack([torch.randint(0, 4, (8, E, A), device=dev), torch.randint(0, w["num_communication"], (8, E, A), device=dev)], -1).to(torch.uint8).contiguous()
    acts_h = [acts_d[i].cpu().pin_memory() for i in range(8)]
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    FL = _cabi.OC_FLAG_AUTO_RESET | _cabi.OC_FLAG_ACTIONS_U8 | _cabi.OC_FLAG_REWARD_PER_ENV
    env.reset()

    def ptrs(base):
        return [C.c_void_p(base + int(o)) for o in (lay.obs_i8, lay.timestep, lay.reward, lay.done)]

    def kernel_dev(i):
        o, t, r, d = ptrs(blk_d.data_ptr())
        lib.check(lib.step_i8(h, C.c_void_p(acts_d[i % 8].data_ptr()), o, t, r, None, d, None, None, FL, st), "step_i8")

    def A_(i):
        kernel_dev(i)
        torch.cuda.synchronize()

    def B_(i):
        blk_h.copy_(blk_d, non_blocking=True)
        torch.cuda.synchronize()

    def C_(i):
        kernel_dev(i)
        blk_h.copy_(blk_d, non_blocking=True)
        torch.cuda.synchronize()

    def F_(i):
        o, t, r, d = ptrs(blk_h.data_ptr())
        lib.check(lib.step_i8(h, C.c_void_p(acts_h[i % 8].data_ptr()), o, t, r, None, d, None, None, FL, st), "step_i8 host")
        torch.cuda.synchronize()

    print("%s: E=%d, block %.2f MB, actions %.2f MB" % (name, E, total / 1e6, acts_h[0].numel() / 1e6))
    # GPU time of the kernel alone (events)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(5):
        kernel_dev(i)
    torch.cuda.synchronize()
    e0.record()
    for i in range(50):
        kernel_dev(i)
    e1.record()
    torch.cuda.synchronize()
    print("   compact step kernel, device time (50 back-to-back launches): %.2f us" % (e0.elapsed_time(e1) * 1e3 / 50))
    for label, fn in (("A kernel alone + sync", A_), ("B pinned D2H copy of the block + sync", B_),
                      ("C kernel + copy + sync", C_), ("F kernel writes all outputs to pinned host memory", F_)):
        m, lo = med(fn)
        print("   %-58s median %7.1f us   min %7.1f us" % (label, m, lo))
    env.close()

    henv = OvercookedHostVecEnv(ns, num_envs=E, seed=1, obs_format="i8", terminal_observations=False)
    hl, hh = henv.lib, henv._handle
    pa = []
    for i in range(8):
        a = henv.pinned_array((E, A, 2), np.uint8)
        a[...] = acts_h[i].numpy()
        pa.append(a)
    henv.reset()
    bp = C.c_void_p(henv._block.ctypes.data)

    def D_(i):
        hl.check(hl.step_host_block(hh, C.c_void_p(pa[i % 8].ctypes.data), bp, None, None, _cabi.OC_FLAG_AUTO_RESET, None), "block")

    def E_(i):
        henv.step(pa[i % 8])

    m, lo = med(D_)
    print("   %-58s median %7.1f us   min %7.1f us" % ("D oc_step_host_block (raw ctypes)", m, lo))
    m, lo = med(E_)
    print("   %-58s median %7.1f us   min %7.1f us" % ("E OvercookedHostVecEnv.step", m, lo))
    enq, wait = [], []
    for i in range(200):
        t0 = time.perf_counter()
        hl.check(hl.step_host_block(hh, C.c_void_p(pa[i % 8].ctypes.data), bp, None, None,
                                    _cabi.OC_FLAG_AUTO_RESET | _cabi.OC_FLAG_NO_SYNC, None), "block")
        t1 = time.perf_counter()
        hl.check(hl.sync(hh, None), "sync")
        t2 = time.perf_counter()
        enq.append(t1 - t0)
        wait.append(t2 - t1)
    print("   %-58s enqueue %6.1f us   wait %7.1f us" % ("G NO_SYNC + oc_sync", statistics.median(enq) * 1e6, statistics.median(wait) * 1e6))
    henv.close()


if __name__ == "__main__":
    main()
