"""world_size-2 gloo test of the domain-decomposition plumbing: partition + halo plan must
reproduce the global SpMV (owned rows from local columns after a ghost update) and the global
dot products (sum of owned parts) on the oracle's Jacobian."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out, cell_type="triangle", partitioner="slab"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import scipy.sparse as sp
    from cfd_hemodynamic_b200.parallel import HaloExchange, HaloExchangeAllGather, Partition, rcb_partition, slab_partition
    from oracle import ns_oracle as O
    from tests import common as T
    gd = 3 if cell_type == "tetrahedron" else 2
    if gd == 3:
        # tetrahedra: the oracle's 3-D Jacobian (cell integrals) on a Kuhn-split box
        from oracle import ns3d_oracle as O3
        from oracle import simplex_oracle as S
        x3, cells3 = O3.unit_cube_tets(3)
        rules = {k: S.tet_gauss_jacobi(d) for k, d in dict(Fu=4, Fp=3, uu=4, up=3, pu=3, pp=2).items()}
        prob = O3.Problem3D(x=x3, cells=cells3, dt=0.01, rho=1.3, mu=0.02, f=np.zeros(3), rules=rules)
        n = prob.n
        rs = np.random.default_rng(5)
        A = O3.assemble_J_raw(prob, rs.standard_normal(4 * n), rs.standard_normal(3 * n)).tocsr()
    else:
        mesh = T.perturbed_square(9, 6, seed=4, cell_type=cell_type)
        prob = T.make_problem(mesh)
        n = prob.n
        u, p, un = T.smooth_fields(prob.x)
        A = O.assemble_J_raw(prob, u, p, un).tocsr()
    owner = slab_partition(prob.x[:, 0], world) if partitioner == "slab" else rcb_partition(prob.x, world)
    part = Partition(prob.x, prob.cells, owner, rank, gdim=gd)
    # local matrix: rows/cols of the local nodes in local [u|p] numbering
    gl = part.glob_nodes
    gdof = np.concatenate([np.stack([gd * gl + k for k in range(gd)], 1).ravel(), gd * n + gl])
    Aloc = A[gdof][:, gdof]
    rng = np.random.default_rng(0)
    xg = rng.standard_normal((gd + 1) * n)
    xl = torch.tensor(xg[gdof].copy())
    ghost_dofs = part.dof_index(np.arange(part.n_owned, part.n_local))
    assert np.array_equal(np.sort(part.dof_index(np.arange(part.n_local))), np.arange((gd + 1) * part.n_local))
    xl[torch.as_tensor(ghost_dofs)] = 0.0            # forget ghost values ...
    HaloExchange(part, torch.device("cpu")).update(xl)   # ... and get them back from the owners
    assert np.allclose(xl.numpy(), xg[gdof], rtol=0, atol=0)
    xl[torch.as_tensor(ghost_dofs)] = -7.0           # same through the single all-gather variant
    HaloExchangeAllGather(part, owner, prob.cells, torch.device("cpu")).update(xl)
    assert np.allclose(xl.numpy(), xg[gdof], rtol=0, atol=0)
    yl = Aloc @ xl.numpy()
    owned_dofs = part.dof_index(np.arange(part.n_owned))
    y_ref = (A @ xg)[gdof][owned_dofs]
    err = np.abs(yl[owned_dofs] - y_ref).max() / np.abs(y_ref).max()
    # the halo plan handed to the library (hemo_comm_set_partition), emulated with P2P on the CPU: pack the send
    # nodes per neighbour as [u (gd each) | p], receive straight into the contiguous ghost slices; with an overlap of
    # several cell layers as well (restricted additive Schwarz)
    from cfd_hemodynamic_b200.parallel import library_halo_plan
    for ov in (1, 3):
        part2 = Partition(prob.x, prob.cells, owner, rank, overlap=ov, gdim=gd)
        peers, send_ptr, send_nodes, recv_ptr = library_halo_plan(part2, owner)
        gl2 = part2.glob_nodes
        gdof2 = np.concatenate([np.stack([gd * gl2 + k for k in range(gd)], 1).ravel(), gd * n + gl2])
        v = torch.tensor(xg[gdof2].copy())
        nl2, no2 = part2.n_local, part2.n_owned
        v[gd * no2:gd * nl2] = 0.0
        v[gd * nl2 + no2:] = 0.0
        ops, bufs = [], []
        for k, q in enumerate(peers.tolist()):
            sn = torch.as_tensor(send_nodes[send_ptr[k]:send_ptr[k + 1]].astype(np.int64))
            su = torch.stack([v[gd * sn + c] for c in range(gd)], 1).reshape(-1).contiguous()
            spp = v[gd * nl2 + sn].contiguous()
            g0, g1 = no2 + int(recv_ptr[k]), no2 + int(recv_ptr[k + 1])
            ru = torch.empty(gd * (g1 - g0), dtype=torch.float64)
            rp = torch.empty(g1 - g0, dtype=torch.float64)
            bufs.append((g0, g1, ru, rp))
            for t_, fn in ((su, dist.isend), (spp, dist.isend), (ru, dist.irecv), (rp, dist.irecv)):
                if t_.numel():
                    ops.append(dist.P2POp(fn, t_, q))
        for w_ in dist.batch_isend_irecv(ops):
            w_.wait()
        for g0, g1, ru, rp in bufs:
            v[gd * g0:gd * g1] = ru
            v[gd * nl2 + g0:gd * nl2 + g1] = rp
        assert np.array_equal(v.numpy(), xg[gdof2]), f"library halo plan, overlap {ov}"
    # distributed dot = allreduce of owned parts
    t = torch.tensor([float(xg[gdof][owned_dofs] @ xg[gdof][owned_dofs])], dtype=torch.float64)
    dist.all_reduce(t)
    out.put((rank, err, float(t.item()), float(xg @ xg), part.n_owned, part.n_local))
    dist.destroy_process_group()


@pytest.mark.parametrize("cell_type", ["triangle", "quadrilateral", "tetrahedron"])
def test_partition_halo_world2(cell_type):
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, cell_type)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, dsum, dref, n_owned, n_local in res:
        assert err < 1e-14, err
        assert abs(dsum - dref) < 1e-12 * dref
        assert n_local > n_owned
    assert res[0][4] + res[1][4] == (64 if cell_type == "tetrahedron" else 70)


def test_partition_halo_world3_two_neighbours():
    """Three slabs: the middle rank has two neighbours and, with an overlap of three cell layers on this narrow mesh, the
    outer ranks hold ghosts owned by the non-adjacent rank as well (the plans of N = 4 / 8 runs have this shape)."""
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 3, port, q, "triangle")) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(3))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, dsum, dref, n_owned, n_local in res:
        assert err < 1e-14, err
        assert abs(dsum - dref) < 1e-12 * dref
        assert n_local > n_owned
    assert sum(r[4] for r in res) == 70


def test_rcb_partition_balance_and_determinism():
    from cfd_hemodynamic_b200.parallel import rcb_partition
    rng = np.random.default_rng(0)
    x = rng.random((1003, 2)) * np.array([2.2, 0.41])
    for parts in (1, 2, 3, 5, 8):
        owner = rcb_partition(x, parts)
        counts = np.bincount(owner, minlength=parts)
        assert counts.sum() == 1003 and counts.max() - counts.min() <= 2 and owner.min() == 0 and owner.max() == parts - 1
        assert np.array_equal(owner, rcb_partition(x, parts))
    # boxes: every part is a union of axis-aligned cuts -> the bounding boxes of two parts overlap in at most a face
    owner = rcb_partition(x, 4)
    boxes = [(x[owner == r].min(axis=0), x[owner == r].max(axis=0)) for r in range(4)]
    for a in range(4):
        for b in range(a + 1, 4):
            lo = np.maximum(boxes[a][0], boxes[b][0])
            hi = np.minimum(boxes[a][1], boxes[b][1])
            assert (hi - lo).min() <= 1e-12


def test_partition_halo_world4_rcb():
    """Four RCB parts of the square: ranks with two or three neighbours (2-D decomposition, the shape an unstructured
    mesh produces); halo plan, library plan with overlap and owned-row SpMV against the global product."""
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 4, port, q, "triangle", "rcb")) for r in range(4)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(4))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, dsum, dref, n_owned, n_local in res:
        assert err < 1e-14, err
        assert abs(dsum - dref) < 1e-12 * dref
        assert n_local > n_owned
    assert sum(r[4] for r in res) == 70
