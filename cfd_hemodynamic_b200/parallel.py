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


# ---------------------------------------------------------------------------
# domain decomposition: vertex ownership + one layer of ghost cells
# (mirrors how DOLFINx distributes a mesh under mpirun, §2.4)
# ---------------------------------------------------------------------------
class Partition:
    """The part of a mesh one rank works on: every cell touching an owned vertex,
    owned vertices first, ghost vertices after; halo plan towards the neighbours."""

    def __init__(self, x, cells, owner, rank: int):
        import numpy as np
        owner = np.asarray(owner)
        cell_owner = owner[cells]                                 # (E, 3)
        mine = (cell_owner == rank).any(axis=1)
        self.rank = rank
        self.cell_glob = np.nonzero(mine)[0]
        lc = cells[self.cell_glob]
        nodes = np.unique(lc)
        owned = nodes[owner[nodes] == rank]
        ghosts = nodes[owner[nodes] != rank]
        self.glob_nodes = np.concatenate([owned, ghosts])
        self.n_owned = int(owned.shape[0])
        self.n_local = int(self.glob_nodes.shape[0])
        g2l = -np.ones(owner.shape[0], dtype=np.int64)
        g2l[self.glob_nodes] = np.arange(self.n_local)
        self.g2l = g2l
        self.cells = g2l[lc].astype(np.int32)
        self.x = np.ascontiguousarray(x[self.glob_nodes])
        self.ghost_mask = np.zeros(self.n_local, dtype=np.uint8)
        self.ghost_mask[self.n_owned:] = 1
        # halo plan: recv = my ghosts owned by q; send = my owned vertices in cells that touch q
        self.neighbors = {}
        for q in np.unique(owner[ghosts]):
            q = int(q)
            recv_g = ghosts[owner[ghosts] == q]
            touch_q = (cell_owner[self.cell_glob] == q).any(axis=1)
            cand = np.unique(lc[touch_q])
            send_g = cand[owner[cand] == rank]
            self.neighbors[q] = (g2l[send_g], g2l[recv_g])

    def dof_index(self, nodes):
        """Indices into a local [u interleaved | p] vector of all dofs of `nodes`."""
        import numpy as np
        nodes = np.asarray(nodes, dtype=np.int64)
        return np.concatenate([2 * nodes, 2 * nodes + 1, 2 * self.n_local + nodes])


class HaloExchange:
    """Forward ghost update (owner -> ghost copies) of local [u|p] vectors:
    the `ghostUpdate(INSERT, FORWARD)` calls of stabilized_schur.py:137-142,168."""

    def __init__(self, part: Partition, device, group=None):
        self.group = group
        self.plan = []
        for q, (send_nodes, recv_nodes) in sorted(part.neighbors.items()):
            si = torch.as_tensor(part.dof_index(send_nodes), dtype=torch.int64, device=device)
            ri = torch.as_tensor(part.dof_index(recv_nodes), dtype=torch.int64, device=device)
            self.plan.append((q, si, ri, torch.empty(si.numel(), dtype=torch.float64, device=device),
                              torch.empty(ri.numel(), dtype=torch.float64, device=device)))
        self.bytes_per_update = sum(8 * (p[1].numel() + p[2].numel()) for p in self.plan)

    def update(self, v: torch.Tensor):
        if not self.plan:
            return
        ops = []
        for q, si, ri, sbuf, rbuf in self.plan:
            torch.index_select(v, 0, si, out=sbuf)
            ops.append(dist.P2POp(dist.isend, sbuf, q, group=self.group))
            ops.append(dist.P2POp(dist.irecv, rbuf, q, group=self.group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for q, si, ri, sbuf, rbuf in self.plan:
            v.index_copy_(0, ri, rbuf)
