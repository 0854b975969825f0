"""Multi-GPU `stabilized_schur`: one mesh partition per GPU (one process per GPU).

Mirrors what the reference gets from `mpirun -n N`: vertex (row) ownership, ghost cells so
that every owned row is assembled locally (no communication in assembly), a forward ghost update before each operator
application and residual evaluation (`ghostUpdate`, stabilized_schur.py:137-142,168), global reductions in the Krylov
and Newton loops (PETSc VecMDot / VecNorm allreduces) and a preconditioner restricted to the partition (PETSc's ASM
sub-solves are rank-local too, :256-267).

Round 2: the exchange steps live in the library (csrc/comm.cu, NCCL loaded there): `hemo_fgmres` runs the
distributed Krylov iteration itself — halo send/recv of the search direction, one ncclAllReduce for the Gram-Schmidt
coefficients and one for the norm, Givens rotations and the convergence flag on the device, the whole iteration
(halo + preconditioner + SpMV + reductions) replayed as one CUDA graph.  This module only builds the partition tables
and sequences the Newton loop; torch.distributed is used for the one-time exchange of the halo plan and the NCCL id.

Preconditioner across ranks: overlapping restricted additive Schwarz with the rank-local block-Schur / AMG
preconditioner, plus — for the pressure operator of the Schur approximation, which needs global coupling — a coarse
space: the level-`coarse_level` operator of the *global* pressure hierarchy, replicated on every rank (a few 10^4
unknowns instead of the full global hierarchy of round 1), applied additively (two-level Schwarz).
"""
from __future__ import annotations

import math
import os

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

from ._lib import Hemo, HemoDiverged, Q_FP, Q_FU, Q_PP, Q_PU, Q_UP, Q_UU
from .fem import amg_setup
from .fem import discretization as D
from .fem import quadrature as Q
from .linear_solver import BlockSchurSolver
from .parallel import Partition, library_halo_plan

BLOCK_DEGREE = {Q_FU: 12, Q_FP: 11, Q_UU: 12, Q_UP: 11, Q_PU: 11, Q_PP: 10}
BLOCK_DEGREE_QUAD = {Q_FU: 22, Q_FP: 20, Q_UU: 22, Q_UP: 20, Q_PU: 20, Q_PP: 18}     # see _stabilized_common.py


def _cell_diameter(x, cells):
    """mesh.h: largest vertex-vertex distance (triangles: longest edge, quadrilaterals: longest diagonal/edge)."""
    X = x[cells]
    nv = cells.shape[1]
    h = np.zeros(cells.shape[0])
    for i in range(nv):
        for j in range(i + 1, nv):
            h = np.maximum(h, np.linalg.norm(X[:, i] - X[:, j], axis=1))
    return h


def _broadcast_nccl_id(rank, group):
    obj = [Hemo.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0, group=group)
    return obj[0]


class DistributedStabilizedSchur:
    def __init__(self, tables: dict, owner: np.ndarray, device_index: int, group=None, verbose=False,
                 coarse_pressure: bool = True, overlap: int | None = None, coarse_level: int = 2):
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.group = group
        self.verbose = verbose and self.rank == 0
        self.variant = tables["variant"]
        par = tables["params"]
        if overlap is None:
            # the velocity block couples over the viscous length sqrt(nu dt); the overlap of the restricted Schwarz
            # solve covers it up to a bound (HEMO_DIST_OVERLAP_MAX layers, default 16)
            hmin = float(_cell_diameter(tables["x"], tables["cells"][:4096]).min()) / math.sqrt(2.0)
            visc = math.sqrt(par["mu"] / par["rho"] * par["dt"]) / max(hmin, 1e-300)
            cap = int(os.environ.get("HEMO_DIST_OVERLAP_MAX", "16"))
            overlap = int(min(max(math.ceil(1.5 * visc), 4), cap))
        self.overlap = int(overlap) if self.world > 1 else 1
        self.part = part = Partition(tables["x"], tables["cells"], owner, self.rank, overlap=self.overlap)
        self.n_global = tables["x"].shape[0]
        self.hemo = hemo = Hemo(device_index)
        dev = hemo.device
        nl = part.n_local
        self.n = nl
        self.N = 3 * nl
        kw = tables["solver_kw"]
        self.snes_rtol, self.snes_atol, self.snes_stol = kw["snes_rtol"], kw["snes_atol"], kw["snes_stol"]
        self.snes_max_it, self.ksp_rtol, self.ksp_max_it = kw["snes_max_it"], kw["ksp_rtol"], kw["ksp_max_it"]
        self.restart = min(kw["ksp_restart"], 100)
        # ---- local tables -------------------------------------------------------
        h = _cell_diameter(part.x, part.cells)
        hemo.set_mesh(torch.from_numpy(part.x).to(dev), torch.from_numpy(part.cells).to(dev),
                      torch.from_numpy(np.ascontiguousarray(h)).to(dev))
        nrowptr, ncol = D.node_graph(part.cells, nl)
        hemo.set_node_graph(torch.from_numpy(nrowptr).to(dev), torch.from_numpy(ncol).to(dev))
        self._quad = quad = part.cells.shape[1] == 4          # Q1 quadrilaterals (tensor-ordered) vs P1 triangles
        for block, deg in (BLOCK_DEGREE_QUAD if quad else BLOCK_DEGREE).items():
            hemo.set_quadrature(block, *(Q.quadrilateral_rule(deg) if quad else Q.triangle_rule(deg)))
        hemo.set_facet_quadrature(*Q.interval_gauss(Q.FACET_POINTS_QUAD if quad else 2))
        hemo.set_params(par["dt"], par["rho"], par["mu"], par["f"], float(np.finfo(np.float64).resolution))
        g2l = part.g2l
        cell_g2l = -np.ones(tables["cells"].shape[0], dtype=np.int64)
        cell_g2l[part.cell_glob] = np.arange(part.cell_glob.shape[0])
        self._outlet = None
        for sid, (pairs, coef) in tables["facet_sets"].items():
            lc = cell_g2l[pairs[:, 0]]
            keep = lc >= 0
            if self.variant in ("pressure_backflow", "velocity_vascular_backflow") and sid == 2:
                # resistance outlet: every overlapping rank sees the outlet cells; the flux is summed over
                # the cells whose first vertex this rank owns so that each facet counts once
                first_owner = owner[tables["cells"][pairs[:, 0], 0]]
                mine = keep & (first_owner == self.rank)
                self._outlet = dict(coef=dict(coef), flux_cells=lc[mine], flux_lf=pairs[mine, 1], **tables["outlet"])
            lp = np.stack([lc[keep], pairs[keep, 1]], axis=1)
            if lp.shape[0] == 0:
                continue
            order = np.argsort(lp[:, 0], kind="stable")
            cells_u, start = np.unique(lp[order, 0], return_index=True)
            mask = np.add.reduceat((1 << lp[order, 1]).astype(np.int32), start).astype(np.int32)
            hemo.set_facet_set(sid, torch.from_numpy(cells_u.astype(np.int32)).to(dev),
                               torch.from_numpy(mask).to(dev), **coef)
        bcs = []
        for block, nodes, values in tables["bcs"]:
            ln = g2l[np.asarray(nodes, dtype=np.int64)]
            sel = ln >= 0
            ln = ln[sel]
            gn = np.asarray(nodes, dtype=np.int64)[sel]
            if block == "u":
                vals = np.zeros(2 * nl)
                vals[2 * ln] = values[2 * gn]
                vals[2 * ln + 1] = values[2 * gn + 1]
            else:
                vals = np.zeros(nl)
                vals[ln] = values[gn]
            bcs.append((block, ln, vals))
        flag, mult, cellflag, g = D.dirichlet_arrays(nl, part.cells, bcs)
        self._has_bc = bool(flag.any())
        if self._has_bc:
            hemo.set_bc(torch.from_numpy(flag).to(dev), torch.from_numpy(mult).to(dev),
                        torch.from_numpy(cellflag).to(dev))
        # the local preconditioner acts on owned + overlap vertices; only vertices with
        # incomplete rows (outermost layer) are held fixed
        hemo.set_pc_mask(torch.from_numpy(part.incomplete_mask).to(dev))
        # ---- communicator + halo plan inside the library ------------------------------------------
        if self.world > 1:
            hemo.comm_init(_broadcast_nccl_id(self.rank, group), self.rank, self.world)
            peers, send_ptr, send_nodes, recv_ptr = library_halo_plan(part, owner, group)
            # no ghost update of the preconditioner input: the Krylov vectors stay valid on the overlap (complete rows)
            hemo.comm_set_partition(part.n_owned, peers, send_ptr, send_nodes, recv_ptr,
                                    ras_overlap=bool(os.environ.get("HEMO_DIST_HALO_PC_INPUT")))
            self.halo_bytes_per_update = int(8 * 3 * (send_ptr[-1] + recv_ptr[-1]))
            self.n_neighbours = int(len(peers))
        else:
            self.halo_bytes_per_update = 0
            self.n_neighbours = 0
        f64 = torch.float64
        self.d_bcval = torch.from_numpy(g).to(dev)
        self.d_x = torch.zeros(self.N, dtype=f64, device=dev)
        self.d_un = torch.zeros(2 * nl, dtype=f64, device=dev)
        self.d_f = torch.zeros(self.N, dtype=f64, device=dev)
        self.d_y = torch.zeros(self.N, dtype=f64, device=dev)
        self.d_w = torch.zeros(self.N, dtype=f64, device=dev)
        self.d_g = torch.zeros(self.N, dtype=f64, device=dev)
        self.d_t = torch.zeros(self.N, dtype=f64, device=dev)
        self.d_trial = torch.zeros(self.N, dtype=f64, device=dev)
        self.d_vals = torch.zeros(hemo.nnz, dtype=f64, device=dev)
        self._red = torch.zeros(8, dtype=f64, device=dev)
        gl = part.glob_nodes
        up = tables["u_prev"].reshape(-1, 2)[gl].reshape(-1)
        self.d_un.copy_(torch.from_numpy(np.ascontiguousarray(up)))
        self.d_x[:2 * nl].copy_(self.d_un)
        self.d_x[2 * nl:].copy_(torch.from_numpy(np.ascontiguousarray(tables["p_prev"][gl])))
        # ---- preconditioner: local block-Schur AMG on owned + overlap ---------------------------------
        hemo.assemble_jacobian(self.d_x, self.d_un, self.d_vals)
        u_nodes = np.nonzero(flag[0:2 * nl:2] | flag[1:2 * nl:2])[0]
        p_nodes = np.nonzero(flag[2 * nl:])[0]
        fixed = np.nonzero(part.incomplete_mask)[0]
        self._nullspace = self._test_nullspace()
        pc_kw = dict(kw["pc_kw"])
        pc_kw["schur_mode"] = "laplace"          # SELFP is single-GPU only
        self.linear = BlockSchurSolver(hemo, nrowptr, ncol, np.union1d(u_nodes, fixed), p_nodes, p_open_nodes=fixed,
                                       dt=par["dt"], rho=par["rho"], mu=par["mu"], restart=self.restart,
                                       max_it=self.ksp_max_it, rtol=self.ksp_rtol,
                                       project_pressure=self._nullspace and self.world == 1, **pc_kw)
        self.hemo_c = None
        self.coarse_n = 0
        if coarse_pressure and self.world > 1:
            self._setup_coarse_pressure(tables, coarse_level, device_index)
        self.linear.setup(self.d_vals)
        self.its_snes = self.its_ksp = 0
        self.reason = 0
        self.n_pressure_global = self.n_global
        self.timers = {}

    # ---- replicated coarse space of the pressure operator ------------------------------------------
    def _setup_coarse_pressure(self, tables, coarse_level, device_index):
        """Global pressure Laplacian of the Schur approximation -> global aggregation hierarchy (host, one-time, the
        same routine the single-GPU solver uses) -> level `coarse_level` becomes the replicated coarse space:
        P0 = P_0 ... P_{k-1} (rows of the local nodes are uploaded), A_c = P0^T L P0 with its own hierarchy below."""
        dev = self.hemo.device
        part = self.part
        xg = np.ascontiguousarray(tables["x"])
        cg = np.ascontiguousarray(tables["cells"], dtype=np.int32)
        ng = xg.shape[0]
        nrp, nc = D.node_graph(cg, ng)
        L = _global_laplacian(xg, cg, nrp, nc)
        pmask = np.zeros(ng, dtype=bool)
        for b, nodes, _ in tables["bcs"]:
            if b == "p":
                pmask[np.asarray(nodes, dtype=np.int64)] = True
        if self.variant != "schur":
            # open (traction) boundaries: Dirichlet rows in the pressure operator of the Schur approximation,
            # as in the single-GPU solver (_stabilized_common.setup)
            from .fem.mesh import Mesh, exterior_facet_indices
            gm = Mesh(xg, cg, cell_type="quadrilateral" if cg.shape[1] == 4 else None)
            ext = exterior_facet_indices(gm.topology)
            bnodes = np.unique(gm.topology.facet_vertices[ext])
            unodes = np.unique(np.concatenate([np.asarray(nodes, dtype=np.int64) for b, nodes, _ in tables["bcs"]
                                               if b == "u"] + [np.zeros(0, np.int64)]))
            pmask[np.setdiff1d(bnodes, unodes)] = True
        lv = amg_setup.build_hierarchy(L, pmask, max_coarse=160)
        k = min(coarse_level, len(lv) - 1)
        if k < 1:
            return
        free = (~pmask).astype(np.float64)
        K = (sp.diags(free) @ L @ sp.diags(free) + sp.diags(pmask.astype(np.float64) * L.diagonal())).tocsr()
        P0 = lv[0]["P"]
        for l in range(1, k):
            P0 = (P0 @ lv[l]["P"]).tocsr()
        Ac = (P0.T @ K @ P0).tocsr()
        Ac.sort_indices()
        pat = lv[k - 1]["C"].tocsr()             # pattern the hierarchy below was built on (superset of Ac's)
        pat.sort_indices()
        ncn = pat.shape[0]
        keys_pat = np.repeat(np.arange(ncn, dtype=np.int64), np.diff(pat.indptr)) * ncn + pat.indices
        keys_a = np.repeat(np.arange(ncn, dtype=np.int64), np.diff(Ac.indptr)) * ncn + Ac.indices
        pos = np.searchsorted(keys_pat, keys_a)
        assert (keys_pat[pos] == keys_a).all()
        vals = np.zeros(pat.nnz)
        vals[pos] = Ac.data
        self.hemo_c = hc = Hemo(device_index)
        hc.set_graph(torch.from_numpy(pat.indptr.astype(np.int32)).to(dev), torch.from_numpy(pat.indices.astype(np.int32)).to(dev))
        for l in range(k, len(lv)):
            d = lv[l]
            hc.amg_set_level(1, l - k, d["P"], d["R"], d["AP"], d["C"])
        hc.amg_finalize(1, len(lv) - k + 1)
        hc.set_solver_opts(**{**self.linear.opts, "project_pressure": 0})
        singular = not pmask.any()
        hc.amg_setup_scalar(torch.from_numpy(vals).to(dev), 1e-8 if singular else 0.0)
        Pl = P0[part.glob_nodes].tocsr()
        Pl.sort_indices()
        Ro = P0[part.glob_nodes[:part.n_owned]].T.tocsr()      # columns = owned local nodes 0 .. n_owned-1
        Ro.sort_indices()
        self.hemo.pc_set_coarse_pressure(hc, Pl, Ro, cycles=self.linear.opts["amg_cycles_p"])
        self.coarse_n = int(ncn)

    # ---- global reductions (device-side sums, NCCL inside the library) ----------------------------------
    def gdot(self, a, b):
        return self.hemo.global_dot(a, b)

    def gnorm(self, a):
        return math.sqrt(max(self.gdot(a, a), 0.0))

    def _allreduce_scalar(self, value: float) -> float:
        self._red[0] = value
        if self.world > 1:
            self.hemo.allreduce(self._red[:1])
        return float(self._red[0].item())

    def _remove_pressure_mean(self, v):
        """MatNullSpaceRemove with the constant-pressure vector over the *global* pressure dofs (no host round trip)."""
        nl, no = self.n, self.part.n_owned
        p = v[2 * nl:]
        self._red[0] = p[:no].sum()
        if self.world > 1:
            self.hemo.allreduce(self._red[:1])
        p -= self._red[0] / self.n_pressure_global

    def _test_nullspace(self):
        nl = self.n
        c = self.d_t
        c.zero_()
        c[2 * nl:] = 1.0 / math.sqrt(self.n_global)
        self.hemo.spmv(self.d_vals, c, self.d_w)
        r = self.gnorm(self.d_w)
        vv = float((self.d_vals * self.d_vals).sum().item())
        scale = math.sqrt(self._allreduce_scalar(vv) if self.world > 1 else vv) / math.sqrt(3 * self.n_global)
        return bool(r < 1e-8 * max(scale, 1e-300))

    # ---- operators ---------------------------------------------------------------------------
    def _residual(self, x, out):
        """x must carry valid ghost values; only the owned entries of `out` are meaningful."""
        self.hemo.assemble_residual(x, self.d_un, self.d_bcval if self._has_bc else None, out)

    def _newton(self):
        hemo = self.hemo
        x, f, y, g, trial = self.d_x, self.d_f, self.d_y, self.d_g, self.d_trial
        self._residual(x, f)
        fnorm = self.gnorm(f)
        if self.verbose:
            print(f"  0 SNES Function norm {fnorm:.12e}")
        ttol = self.snes_rtol * fnorm
        lin_its = 0
        if fnorm < self.snes_atol:
            return 0, 0, 2
        for it in range(self.snes_max_it):
            hemo.assemble_jacobian(x, self.d_un, self.d_vals)
            if it == 0:
                self.linear.setup(self.d_vals)
            try:
                kits, _ = hemo.fgmres(self.d_vals, f, y)          # distributed Krylov iteration inside the library
            except HemoDiverged:
                return it, lin_its, -3
            lin_its += kits
            hemo.spmv(self.d_vals, y, self.d_t)                   # y returns with valid ghosts
            slope = self.gdot(f, self.d_t)
            slope = -abs(slope) if slope != 0.0 else -1.0
            alpha, lam = 1e-4, 1.0
            f2 = 0.5 * fnorm * fnorm
            lam_prev = g_prev = None
            accepted = False
            for _ in range(40):
                trial.copy_(x)
                hemo.axpy(-lam, y, trial)                 # y and x both carry valid ghosts
                self._residual(trial, g)
                gnorm = self.gnorm(g)
                g2 = 0.5 * gnorm * gnorm
                if math.isfinite(gnorm) and g2 <= f2 + lam * alpha * slope:
                    accepted = True
                    break
                if not math.isfinite(gnorm):
                    lam_new = 0.5 * lam
                elif lam_prev is None:
                    lam_new = -slope / (2.0 * (g2 - f2 - slope))
                else:
                    t1 = g2 - f2 - lam * slope
                    t2 = g_prev - f2 - lam_prev * slope
                    a = (t1 / lam ** 2 - t2 / lam_prev ** 2) / (lam - lam_prev)
                    b = (-lam_prev * t1 / lam ** 2 + lam * t2 / lam_prev ** 2) / (lam - lam_prev)
                    disc = max(b * b - 3.0 * a * slope, 0.0)
                    lam_new = -slope / (2.0 * b) if a == 0.0 else (-b + math.sqrt(disc)) / (3.0 * a)
                lam_new = min(max(lam_new, 0.1 * lam), 0.5 * lam)
                lam_prev, g_prev = lam, g2
                lam = lam_new
            if not accepted:
                return it + 1, lin_its, -6
            ynorm = lam * self.gnorm(y)
            x.copy_(trial)
            f.copy_(g)
            fnorm = gnorm
            if self.verbose:
                print(f"  {it + 1} SNES Function norm {fnorm:.12e} ({kits} FGMRES its)")
            if fnorm < self.snes_atol:
                return it + 1, lin_its, 2
            if fnorm <= ttol:
                return it + 1, lin_its, 3
            if ynorm < self.snes_stol * self.gnorm(x):
                return it + 1, lin_its, 4
        return self.snes_max_it, lin_its, -5

    def _update_outlet_pressure(self):
        """p_c <- alpha R |Q| + (1 - alpha) p_c with Q = global int u_prev.n ds_out
        (stabilized_schur_pressure_backflow.py:383-396): local flux over owned outlet cells + allreduce."""
        o = self._outlet
        part = self.part
        cells = part.cells[o["flux_cells"]]
        if cells.shape[0]:
            X = part.x[cells]
            lf = o["flux_lf"]
            ar = np.arange(cells.shape[0])
            if cells.shape[1] == 4:
                from .fem.mesh import QUAD_FACETS
                fv = np.array(QUAD_FACETS)
                inside = X.mean(axis=1)
            else:
                fv = np.array([[1, 2], [0, 2], [0, 1]])
                inside = X[ar, lf]
            va, vb = fv[lf, 0], fv[lf, 1]
            t = X[ar, vb] - X[ar, va]
            nrm = np.stack([t[:, 1], -t[:, 0]], axis=1)
            nrm *= np.sign(np.einsum("ei,ei->e", nrm, 0.5 * (X[ar, va] + X[ar, vb]) - inside))[:, None]
            nodes = np.unique(cells)
            un = np.zeros((self.n, 2))
            idx = torch.from_numpy(np.concatenate([2 * nodes, 2 * nodes + 1])).to(self.hemo.device)
            vals = self.d_un.index_select(0, idx).cpu().numpy()
            un[nodes, 0] = vals[:nodes.shape[0]]
            un[nodes, 1] = vals[nodes.shape[0]:]
            U = un[cells]
            q_loc = float(np.sum(np.einsum("ei,ei->e", 0.5 * (U[ar, va] + U[ar, vb]), nrm)))
        else:
            q_loc = 0.0
        q = self._allreduce_scalar(q_loc)
        o["p_c"] = o["alpha_damping"] * o["R_resistance"] * abs(q) + (1.0 - o["alpha_damping"]) * o["p_c"]
        coef = dict(o["coef"])
        coef["pconst"] = 0.5 * (sum(o["p_c_frozen"]) + o["p_c"])
        self.hemo.set_facet_coef(2, **coef)

    def step_device(self):
        """One time step, everything resident on the GPUs; u_prev <- u_sol on the device."""
        self._remove_pressure_mean(self.d_x)      # nullsp.remove(x_n), unconditional (:319)
        self.its_snes, self.its_ksp, self.reason = self._newton()
        if self.reason < 0:
            raise RuntimeError(f"Did not converge, reason: {self.reason}.")
        if self._outlet is not None:
            self._update_outlet_pressure()        # Q from the old u_prev (one-step lag, scenario.py:306)
        self.d_un.copy_(self.d_x[:2 * self.n])

    def launches(self) -> int:
        return self.hemo.launches          # the replicated coarse context's launches are counted by the library

    def comm_summary(self) -> dict:
        out = {"decomposition": f"{self.world} vertex-owned x-slabs", "overlap_cell_layers": self.overlap,
               "owned_nodes": self.part.n_owned, "local_nodes": self.part.n_local, "neighbours": self.n_neighbours,
               "halo_bytes_per_update": self.halo_bytes_per_update, "coarse_pressure_space": self.coarse_n,
               "krylov": "hemo_fgmres in the library: NCCL send/recv halo + 2 allreduces per iteration, device-side Givens, "
                         "one CUDA graph per iteration"}
        if self.world > 1:
            out.update({k: v for k, v in self.hemo.comm_info().items() if k in ("nccl_version", "halo_updates", "allreduces")})
        return out

    def gather_solution(self):
        """(u, p) in global numbering on every rank (test / output helper)."""
        part = self.part
        no = part.n_owned
        xl = self.d_x.cpu().numpy()
        u_own = xl[:2 * self.n].reshape(-1, 2)[:no]
        p_own = xl[2 * self.n:][:no]
        objs = [None] * self.world
        dist.all_gather_object(objs, (part.glob_nodes[:no], u_own, p_own), group=self.group)
        u = np.zeros((self.n_global, 2))
        p = np.zeros(self.n_global)
        for nodes, uu, pp in objs:
            u[nodes] = uu
            p[nodes] = pp
        return u.reshape(-1), p


def _global_laplacian(x, cells, nrowptr, ncol):
    """P1 / Q1 stiffness matrix of the pressure space on the node graph (host, setup only)."""
    n = x.shape[0]
    if cells.shape[1] == 3:
        X = x[cells]
        J = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]], axis=2)
        det = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
        Jinv = np.empty_like(J)
        Jinv[:, 0, 0], Jinv[:, 0, 1] = J[:, 1, 1] / det, -J[:, 0, 1] / det
        Jinv[:, 1, 0], Jinv[:, 1, 1] = -J[:, 1, 0] / det, J[:, 0, 0] / det
        ghat = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])
        dphi = np.einsum("aj,eji->eai", ghat, Jinv)
        Ke = 0.5 * np.abs(det)[:, None, None] * np.einsum("eai,ebi->eab", dphi, dphi)
    else:
        # Q1: 2x2 Gauss on the bilinear map
        gp = np.array([0.5 - 0.5 / math.sqrt(3.0), 0.5 + 0.5 / math.sqrt(3.0)])
        X = x[cells]
        Ke = np.zeros((cells.shape[0], 4, 4))
        for xi in gp:
            for eta in gp:
                dN = np.array([[-(1 - eta), -(1 - xi)], [(1 - eta), -xi], [-eta, (1 - xi)], [eta, xi]])
                J = np.einsum("ai,eaj->eji", dN, X)           # J[e, j, i] = dx_j / dxi_i
                det = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
                Jinv = np.empty_like(J)
                Jinv[:, 0, 0], Jinv[:, 0, 1] = J[:, 1, 1] / det, -J[:, 0, 1] / det
                Jinv[:, 1, 0], Jinv[:, 1, 1] = -J[:, 1, 0] / det, J[:, 0, 0] / det
                dphi = np.einsum("ai,eij->eaj", dN, Jinv)
                Ke += 0.25 * np.abs(det)[:, None, None] * np.einsum("eai,ebi->eab", dphi, dphi)
    nv = cells.shape[1]
    rows = np.repeat(cells, nv, axis=1).reshape(-1)
    cols = np.tile(cells, (1, nv)).reshape(-1)
    L = sp.coo_matrix((Ke.reshape(-1), (rows, cols)), shape=(n, n)).tocsr()
    L.sort_indices()
    return L
