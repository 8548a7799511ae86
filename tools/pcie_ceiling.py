"""The host side's ceiling for the host-buffer path, measured on the box itself: every rank (one per GPU, under
torchrun) copies blocks the size of one cfg2 step's result (6.49 MB) from its GPU into page-locked host memory, all
ranks at the same time, and the aggregate rate is what N GPUs can deliver to the host together -- the number the e2e
line of bench.py cannot exceed (x 131,072 agent-steps per 6.49 MB).

    python tools/pcie_ceiling.py                                   # one GPU
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_ceiling.py
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_comm_b200.sharding import bind_cpu_to_device  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    aff = bind_cpu_to_device(local) if world > 1 else False
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = {}
    for label, nbytes in (("cfg2 step block 6.49 MB", 6488064), ("64 MiB", 64 << 20)):
        d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        for direction in ("d2h", "h2d"):
            src, dst = (d, h) if direction == "d2h" else (h, d)
            for _ in range(5):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            n = 0
            t0 = time.perf_counter()
            while time.perf_counter() - t0 < 0.5:
                for _ in range(20):
                    dst.copy_(src, non_blocking=True)
                torch.cuda.synchronize()
                n += 20
            dt = time.perf_counter() - t0
            rate = torch.tensor([nbytes * n / dt / 1e9], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(rate)
            out["%s %s" % (label, direction)] = {"aggregate_GBps": float(rate.item()), "per_gpu_GBps": float(rate.item()) / world}
    if rank == 0:
        blk = out["cfg2 step block 6.49 MB d2h"]["aggregate_GBps"]
        print(json.dumps({"n_gpus": world, "cpu_affinity": bool(aff), "copies": out,
                          "e2e_ceiling_agent_steps_per_s": blk * 1e9 / 6488064 * 131072,
                          "note": "pinned cudaMemcpyAsync, all ranks concurrently, back-to-back copies (no kernel, no sync per copy)"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
