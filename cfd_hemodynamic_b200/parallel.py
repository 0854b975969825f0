"""`mesh.comm` stand-in on torch.distributed (one process per GPU).

The reference's host layer uses an mpi4py communicator only for scalar
allreduce / barrier / gather in the time loop (src/scenario.py:206,273-280,
316-319; src/solvers/stabilized_schur_pressure_backflow.py:205,385).  This
class provides those calls over torch.distributed (NCCL on GPUs, gloo in the
CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist

SUM, MAX, MIN = "sum", "max", "min"
_OPS = {SUM: dist.ReduceOp.SUM, MAX: dist.ReduceOp.MAX, MIN: dist.ReduceOp.MIN, None: dist.ReduceOp.SUM}


class TorchComm:
    def __init__(self, device: torch.device | None = None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.rank = dist.get_rank()
        self.size = dist.get_world_size()
        self.device = device or torch.device("cpu")

    def barrier(self):
        dist.barrier()

    Barrier = barrier

    def allreduce(self, value, op=SUM):
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device)
        dist.all_reduce(t, op=_OPS[op])
        return float(t.item())

    def bcast(self, value, root=0):
        obj = [value]
        dist.broadcast_object_list(obj, src=root)
        return obj[0]

    def gather(self, value, root=0):
        out = [None] * self.size if self.rank == root else None
        dist.gather_object(value, out, dst=root)
        return out


def slab_partition(x_coord, n_parts: int):
    """Contiguous x-slabs with equal vertex counts: returns the
    owner rank of every vertex.  Used by the partition-invariance tests; the
    multi-GPU solve itself is replica-parallel in this round."""
    import numpy as np
    order = np.argsort(x_coord, kind="stable")
    owner = np.empty(x_coord.shape[0], dtype=np.int32)
    bounds = np.linspace(0, x_coord.shape[0], n_parts + 1).astype(np.int64)
    for r in range(n_parts):
        owner[order[bounds[r]:bounds[r + 1]]] = r
    return owner
