"""world_size-2 gloo test of the N>1 host path (torch.distributed plumbing used by
bench.py under torchrun: comm wrapper, max-over-ranks timing, slab partition)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cfd_hemodynamic_b200.parallel import MAX, SUM, TorchComm
    comm = TorchComm()
    assert comm.size == world and comm.rank == rank
    s = comm.allreduce(rank + 1.0, op=SUM)
    m = comm.allreduce(10.0 * (rank + 1), op=MAX)
    b = comm.bcast({"nx": 7} if rank == 0 else None, root=0)
    g = comm.gather(rank * 2, root=0)
    comm.barrier()
    # bench-style aggregate: value = world * dofs * steps / max-over-ranks(time)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out.put((rank, s, m, b["nx"], g, float(t.item())))
    dist.destroy_process_group()


def test_gloo_world2():
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1:4] == (3.0, 20.0, 7) and res[1][1:4] == (3.0, 20.0, 7)
    assert res[0][4] == [0, 2] and res[1][4] is None
    assert res[0][5] == 2.0 and res[1][5] == 2.0


def test_slab_partition_balanced():
    from cfd_hemodynamic_b200.parallel import slab_partition
    x = np.random.default_rng(0).random(1001)
    owner = slab_partition(x, 4)
    counts = np.bincount(owner, minlength=4)
    assert counts.max() - counts.min() <= 1
    for r in range(3):
        assert x[owner == r].max() <= x[owner == r + 1].min()
