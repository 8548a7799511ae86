"""Data-parallel training on CPU (world_size 2, gloo): each rank steps its own env shard (the emulated env), the
learners start from rank 0's weights and average their gradients every minibatch, the episode statistics are
summed over the ranks.  On the GPU box the same code runs over NCCL (`torchrun ... train_ppo.py`)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _factory(ns, args):
    from gym_comm_b200.vec_env import OvercookedVecEnv
    from tests.parity_util import EmuVecEnv, emu_library
    return EmuVecEnv(ns, num_envs=args.envs, device="cpu", seed=args.env_seed, auto_reset=True, lib=emu_library())


def _worker(rank, world, port, flags, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gym_comm_b200.ppo import average_gradients, broadcast_parameters
    import train_ppo

    # the collective itself: gradients of different data -> their mean on every rank; weights -> rank 0's
    torch.manual_seed(rank)
    lin = torch.nn.Linear(3, 2)
    broadcast_parameters(lin)
    w0 = [torch.zeros_like(lin.weight) for _ in range(world)]
    dist.all_gather(w0, lin.weight.detach())
    lin(torch.full((1, 3), float(rank + 1))).sum().backward()
    average_gradients(lin)
    ok = all(torch.equal(w, w0[0]) for w in w0) and torch.allclose(lin.weight.grad, torch.full((2, 3), (1.0 + world) / 2.0))

    learners = []
    hist = train_ppo.main(["--envs", "8", "--n-steps", "5", "--iters", "2", "--log-every", "1", "--batch-size", "20",
                           "--max-num-timesteps", "4", "--epochs", "2", "--device", "cpu", "--seed", "3"] + flags,
                          env_factory=_factory, learners_out=learners)
    flats = []
    for m in learners:
        flat = torch.cat([p.detach().reshape(-1) for p in m.policy.parameters()])
        g = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(g, flat)
        flats.append(g)
    dist.barrier()
    if rank == 0:
        out.put(dict(ok=bool(ok), hist=hist, same=[all(torch.equal(x, g[0]) for x in g) for g in flats],
                     moved=[float(g[0].abs().sum()) for g in flats]))
    dist.destroy_process_group()


@pytest.mark.parametrize("flags", [[], ["--recurrent", "--lstm-hidden", "8"]])
def test_two_rank_training_keeps_the_learners_identical(flags):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, flags, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["ok"]
    assert res["same"] == [True, True]                       # ego and partner: bit-identical weights on both ranks
    h = res["hist"]
    assert len(h) == 2 and h[-1]["world"] == 2 and h[-1]["env_steps"] == 2 * 5 * 8 * 2
    assert h[0]["episodes"] + h[1]["episodes"] == 2 * 8 * 2   # 10 steps of T = 4: two episodes per env, both shards counted
