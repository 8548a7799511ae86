"""N>1 host logic on CPU: two gloo ranks agree on a disjoint cover of the env index range, get
distinct seeds, and reduce timings with max-over-ranks (what bench.py does over NCCL)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gym_comm_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert sharding.rank_world() == (rank, world, rank)
    lo, hi = sharding.shard_range(total, rank, world)
    gathered = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, torch.tensor([lo, hi, sharding.shard_seed(5, rank) & 0x7FFFFFFF]))
    slow = sharding.max_over_ranks(1.0 + rank)
    total_steps = sharding.sum_over_ranks(float(hi - lo))
    dist.barrier()
    if rank == 0:
        out.put(([g.tolist() for g in gathered], slow, total_steps))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    world, total = 2, 65537
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    ranges, slow, total_steps = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ranges[0][0] == 0 and ranges[0][1] == ranges[1][0] and ranges[1][1] == total
    assert abs((ranges[0][1] - ranges[0][0]) - (ranges[1][1] - ranges[1][0])) <= 1
    assert ranges[0][2] != ranges[1][2]
    assert slow == 2.0 and total_steps == float(total)


def test_shard_range_properties():
    for total in (0, 1, 7, 65536, 1048576 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_bind_cpu_to_device_is_safe_without_a_gpu(monkeypatch):
    """No NVML / no GPU: the helper reports False and leaves the affinity mask alone."""
    import os
    from gym_comm_b200.sharding import bind_cpu_to_device
    before = os.sched_getaffinity(0)
    monkeypatch.setenv("OC_NO_AFFINITY", "1")
    assert bind_cpu_to_device(0) is False
    monkeypatch.delenv("OC_NO_AFFINITY")
    import torch
    if not torch.cuda.is_available():
        assert bind_cpu_to_device(0) is False
        assert os.sched_getaffinity(0) == before
