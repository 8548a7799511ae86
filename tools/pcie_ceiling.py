"""The host side's ceiling for the host-buffer path, measured on the box itself (no env, no kernel of ours): every
rank (one per GPU, under torchrun) copies blocks the size of one cfg2 step's result (6.49 MB) from its GPU into
page-locked host memory with plain cudaMemcpyAsync, all ranks at the same time.  Two patterns:

  stream   copies back to back, one synchronise per 20 (what the link and the host memory system can sustain)
  stepwise one copy, one cudaStreamSynchronize, repeat (the pattern of a synchronous `step()`: adds the per-copy
           launch + completion latency)

The aggregate rate is what N GPUs can deliver to the host together; x 131,072 agent-steps per 6.49 MB it is the number
the `e2e` line of bench.py cannot exceed.

    python tools/pcie_ceiling.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_ceiling.py
"""
import json
import os
import sys
import time
import warnings

import torch
import torch.distributed as dist

warnings.simplefilter("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_comm_b200.sharding import bind_cpu_to_device  # noqa: E402

BLOCK = 6488064          # oc_host_block_layout(cfg2).total_bytes: obs_i8 | timestep | reward | done of 65,536 envs


def main():
    from cuda.bindings import runtime as rt
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    aff = bind_cpu_to_device(local) if world > 1 else False       # pinned buffers land on the GPU's own NUMA node
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rt.cudaSetDevice(local)
    out = {}
    for label, nbytes in (("block 6.49 MB", BLOCK), ("64 MiB", 64 << 20)):
        err, dptr = rt.cudaMalloc(nbytes)
        err2, hptr = rt.cudaHostAlloc(nbytes, rt.cudaHostAllocDefault)
        err3, st = rt.cudaStreamCreate()
        assert int(err) == 0 and int(err2) == 0 and int(err3) == 0
        for direction, kind in (("d2h", rt.cudaMemcpyKind.cudaMemcpyDeviceToHost), ("h2d", rt.cudaMemcpyKind.cudaMemcpyHostToDevice)):
            dst, src = (hptr, dptr) if direction == "d2h" else (dptr, hptr)
            for pattern, per_sync in (("stream", 20), ("stepwise", 1)):
                for _ in range(5):
                    rt.cudaMemcpyAsync(dst, src, nbytes, kind, st)
                rt.cudaStreamSynchronize(st)
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                n = 0
                t0 = time.perf_counter()
                while time.perf_counter() - t0 < 0.4:
                    for _ in range(per_sync):
                        rt.cudaMemcpyAsync(dst, src, nbytes, kind, st)
                    rt.cudaStreamSynchronize(st)
                    n += per_sync
                dt = time.perf_counter() - t0
                rate = torch.tensor([nbytes * n / dt / 1e9], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(rate)
                out["%s %s %s" % (label, direction, pattern)] = {"aggregate_GBps": round(float(rate.item()), 2),
                                                                 "per_gpu_GBps": round(float(rate.item()) / world, 2),
                                                                 "us_per_copy": round(dt / n * 1e6, 1)}
        rt.cudaFree(dptr)
        rt.cudaFreeHost(hptr)
    if rank == 0:
        a = out["block 6.49 MB d2h stream"]["aggregate_GBps"]
        b = out["block 6.49 MB d2h stepwise"]["aggregate_GBps"]
        print(json.dumps({"n_gpus": world, "cpu_affinity": bool(aff), "copies": out,
                          "e2e_ceiling_agent_steps_per_s": {"stream": a * 1e9 / BLOCK * 131072, "stepwise": b * 1e9 / BLOCK * 131072},
                          "note": "pinned cudaMemcpyAsync of one cfg2 step's result block, all ranks concurrently; no kernel, "
                                  "no Python env; e2e of bench.py additionally runs the step kernel and reads the actions"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
