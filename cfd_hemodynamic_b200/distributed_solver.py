"""Multi-GPU `stabilized_schur`: one mesh partition per GPU (one process per GPU).

Mirrors what the reference gets from `mpirun -n N`: vertex
(row) ownership, one layer of ghost cells so every owned row is assembled locally
(no communication in assembly), a forward ghost update before each operator
application and residual evaluation (`ghostUpdate`, stabilized_schur.py:137-142,168),
global reductions in the Krylov and Newton loops (PETSc VecMDot/VecNorm allreduces),
and a preconditioner restricted to the partition (PETSc's ASM sub-solves are rank-local
too, :256-267).  NCCL (torch.distributed) is used only for the halo exchange and the
allreduces; every kernel is libhemo_sm100.so.

Supported: variants whose boundary-term coefficients are static ("schur", "backflow").
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.distributed as dist

from ._lib import Hemo, Q_FP, Q_FU, Q_PP, Q_PU, Q_UP, Q_UU
from .fem import discretization as D
from .fem import quadrature as Q
from .linear_solver import BlockSchurSolver
from .parallel import HaloExchange, HaloExchangeAllGather, Partition

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


class DistributedStabilizedSchur:
    def __init__(self, tables: dict, owner: np.ndarray, device_index: int, group=None, verbose=False,
                 global_pressure: bool = True, overlap: int | None = None):
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.group = group
        self.verbose = verbose and self.rank == 0
        self.variant = tables["variant"]
        if overlap is None:
            # the velocity block couples over the viscous length sqrt(nu dt); the overlap of the
            # restricted Schwarz solve has to cover it or the iteration count grows with N
            par0 = tables["params"]
            hmin = float(_cell_diameter(tables["x"], tables["cells"][:4096]).min()) / math.sqrt(2.0)
            visc = math.sqrt(par0["mu"] / par0["rho"] * par0["dt"]) / max(hmin, 1e-300)
            overlap = int(min(max(math.ceil(1.5 * visc), 4), 32))
        self.overlap = int(overlap) if dist.get_world_size(group) > 1 else 1
        self.part = part = Partition(tables["x"], tables["cells"], owner, self.rank, overlap=self.overlap)
        self.n_global = tables["x"].shape[0]
        self.hemo = hemo = Hemo(device_index)
        dev = hemo.device
        nl = part.n_local
        self.n = nl
        self.N = 3 * nl
        par = tables["params"]
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
        self.ghost_mask = torch.from_numpy(part.ghost_mask).to(dev)
        # the local preconditioner acts on owned + overlap vertices; only vertices with
        # incomplete rows (outermost layer) are held fixed
        hemo.set_pc_mask(torch.from_numpy(part.incomplete_mask).to(dev))
        self.halo = HaloExchangeAllGather(part, owner, tables["cells"], dev, group)
        f64 = torch.float64
        self.d_bcval = torch.from_numpy(g).to(dev)
        self.d_x = torch.zeros(self.N, dtype=f64, device=dev)
        self.d_un = torch.zeros(2 * nl, dtype=f64, device=dev)
        self.d_f = torch.zeros(self.N, dtype=f64, device=dev)
        self.d_y = torch.zeros(self.N, dtype=f64, device=dev)
        self.d_w = torch.zeros(self.N, dtype=f64, device=dev)
        self.d_g = torch.zeros(self.N, dtype=f64, device=dev)
        self.d_t = torch.zeros(self.N, dtype=f64, device=dev)
        self.d_vals = torch.zeros(hemo.nnz, dtype=f64, device=dev)
        self.ldv = (self.N + 31) // 32 * 32
        self.V = torch.zeros((self.restart + 1) * self.ldv, dtype=f64, device=dev)
        self.Z = torch.zeros(self.restart * self.ldv, dtype=f64, device=dev)
        self._red = torch.zeros(self.restart + 4, dtype=f64, device=dev)
        gl = part.glob_nodes
        up = tables["u_prev"].reshape(-1, 2)[gl].reshape(-1)
        self.d_un.copy_(torch.from_numpy(np.ascontiguousarray(up)))
        self.d_x[:2 * nl].copy_(self.d_un)
        self.d_x[2 * nl:].copy_(torch.from_numpy(np.ascontiguousarray(tables["p_prev"][gl])))
        # ---- preconditioner: local block-Schur AMG, ghosts excluded -----------------------
        hemo.assemble_jacobian(self.d_x, self.d_un, self.d_vals)
        u_nodes = np.nonzero(flag[0:2 * nl:2] | flag[1:2 * nl:2])[0]
        p_nodes = np.nonzero(flag[2 * nl:])[0]
        ghosts = np.nonzero(part.incomplete_mask)[0]
        p_open = ghosts
        self._nullspace = self._test_nullspace()
        self.linear = BlockSchurSolver(hemo, nrowptr, ncol, np.union1d(u_nodes, ghosts), p_nodes, p_open_nodes=p_open,
                                       dt=par["dt"], rho=par["rho"], mu=par["mu"], restart=self.restart,
                                       max_it=self.ksp_max_it, rtol=self.ksp_rtol, project_pressure=False,
                                       **kw["pc_kw"])
        self.linear.setup(self.d_vals)
        self.its_snes = self.its_ksp = 0
        self.reason = 0
        import os
        self._profile = bool(os.environ.get("HEMO_DIST_PROFILE"))
        self.timers = {}
        self._last_key = None
        self._last_t = 0.0
        self.n_pressure_global = self.n_global
        self.global_pressure = bool(global_pressure) and self.world > 1
        if self.global_pressure:
            self._setup_global_pressure(tables, owner, device_index)

    # ---- replicated global pressure solve -------------------------------------------------
    def _setup_global_pressure(self, tables, owner, device_index):
        """The Schur-complement approximation needs L_p^-1 with *global* coupling (a
        partition-local Laplacian solve loses the low modes and the Krylov iteration count
        grows ~4x).  The scalar pressure problem is small next to the 3x3-block Jacobian, so
        every rank keeps the global pressure Laplacian hierarchy and applies the V-cycle
        redundantly; only the pressure residual is all-gathered over NVLink."""
        from .fem import amg_setup
        import scipy.sparse as sp
        dev = self.hemo.device
        xg = np.ascontiguousarray(tables["x"])
        cg = np.ascontiguousarray(tables["cells"], dtype=np.int32)
        ng = xg.shape[0]
        self.hemo_p = hp = Hemo(device_index)
        hp.set_mesh(torch.from_numpy(xg).to(dev), torch.from_numpy(cg).to(dev),
                    torch.from_numpy(np.ascontiguousarray(_cell_diameter(xg, cg))).to(dev))
        nrp, nc = D.node_graph(cg, ng)
        hp.set_node_graph(torch.from_numpy(nrp).to(dev), torch.from_numpy(nc).to(dev))
        pbc = [(b, nodes, vals) for b, nodes, vals in tables["bcs"] if b == "p"]
        pmask = np.zeros(ng, dtype=bool)
        if pbc:
            flag, mult, cellflag, _ = D.dirichlet_arrays(ng, cg, pbc)
            hp.set_bc(torch.from_numpy(flag).to(dev), torch.from_numpy(mult).to(dev), torch.from_numpy(cellflag).to(dev))
            pmask = flag[2 * ng:].astype(bool)
        self.g_lap, self.g_mass = hp.assemble_laplace_mass()
        lap_host = self.g_lap.cpu().numpy()
        if self.variant != "schur":
            # open (traction) boundaries: Dirichlet rows in the pressure operator of the Schur
            # approximation, as in the single-GPU solver (_stabilized_common.setup)
            from .fem.mesh import Mesh, exterior_facet_indices
            gm = Mesh(xg, cg, cell_type="quadrilateral" if cg.shape[1] == 4 else None)
            ext = exterior_facet_indices(gm.topology)
            bnodes = np.unique(gm.topology.facet_vertices[ext])
            unodes = np.unique(np.concatenate([np.asarray(nodes, dtype=np.int64) for b, nodes, _ in tables["bcs"]
                                               if b == "u"] + [np.zeros(0, np.int64)]))
            openn = np.setdiff1d(bnodes, unodes)
            omask = np.zeros(ng, dtype=bool)
            omask[openn] = True
            rows = np.repeat(np.arange(ng), np.diff(nrp))
            hit = omask[rows] | omask[nc]
            lap_host = np.where(hit, (rows == nc).astype(np.float64), lap_host)
            self.g_lap = torch.from_numpy(np.ascontiguousarray(lap_host)).to(dev)
            pmask = pmask | omask
        L = sp.csr_matrix((lap_host, nc, nrp), shape=(ng, ng))
        lv = amg_setup.build_hierarchy(L, pmask, max_coarse=160)
        for l, d in enumerate(lv):
            hp.amg_set_level(1, l, d["P"], d["R"], d["AP"], d["C"])
        hp.amg_finalize(1, len(lv) + 1)
        hp.set_solver_opts(**{**self.linear.opts, "project_pressure": 0})
        hp.amg_setup_scalar(self.g_lap, 1e-8 if self._nullspace else 0.0)
        self.hemo.set_external_schur(True)
        self.hemo.use_graph(True)
        self.linear._first = False
        self.hemo.pc_setup(self.d_vals)            # re-capture the graph without the local pressure part
        # gather plan: owned pressure values of every rank -> global vector
        counts = [int((owner == q).sum()) for q in range(self.world)]
        self._gmax = max(counts)
        idx = torch.full((self.world, self._gmax), ng, dtype=torch.int64)      # padding -> dummy slot ng
        for q in range(self.world):
            own_q = np.nonzero(owner == q)[0]
            idx[q, :counts[q]] = torch.from_numpy(own_q)
        self._gidx = idx.reshape(-1).to(dev)
        self._gsend = torch.zeros(self._gmax, dtype=torch.float64, device=dev)
        self._grecv = torch.zeros(self.world * self._gmax, dtype=torch.float64, device=dev)
        self._rp_g = torch.zeros(ng + 1, dtype=torch.float64, device=dev)
        self._q_g = torch.zeros(ng, dtype=torch.float64, device=dev)
        self._loc_nodes = torch.from_numpy(self.part.glob_nodes).to(dev)
        self._pbc_g = torch.from_numpy(np.nonzero(pmask)[0]).to(dev) if pmask.any() else None
        self._cm = self.linear.opts["schur_mass_coef"]
        self._cl = self.linear.opts["schur_lap_coef"]
        self._cyc_p = self.linear.opts["amg_cycles_p"]

    def _global_schur(self, r, z):
        """z_p (all local nodes, ghosts included) = c_m r_p / m + c_L L_glob^-1 r_p."""
        nl, no, ng = self.n, self.part.n_owned, self.n_global
        self._gsend[:no].copy_(r[2 * nl:2 * nl + no])
        dist.all_gather_into_tensor(self._grecv, self._gsend, group=self.group)
        self._rp_g.index_copy_(0, self._gidx, self._grecv)
        rp = self._rp_g[:ng]
        if self._nullspace:
            rp -= rp.mean()
        self.hemo_p.amg_apply(1, rp, self._q_g, self._cyc_p)
        zg = self._q_g
        zg.mul_(self._cl).add_(rp / self.g_mass, alpha=self._cm)
        if self._pbc_g is not None:
            zg.index_copy_(0, self._pbc_g, rp.index_select(0, self._pbc_g))
        if self._nullspace:
            zg -= zg.mean()
        torch.index_select(zg, 0, self._loc_nodes, out=z[2 * nl:])

    # ---- global reductions ---------------------------------------------------------------
    def _allreduce(self, values):
        k = len(values)
        t = self._red[:k]
        t.copy_(torch.as_tensor(values, dtype=torch.float64))
        dist.all_reduce(t, group=self.group)
        return t.cpu().numpy().copy()

    def gdot(self, a, b):
        """Global dot product; at least one operand must have zero ghost entries."""
        return float(self._allreduce([self.hemo.dot(a, b)])[0])

    def gnorm(self, a_zero_ghosts):
        return math.sqrt(max(self.gdot(a_zero_ghosts, a_zero_ghosts), 0.0))

    def _remove_pressure_mean(self, v):
        """MatNullSpaceRemove with the constant-pressure vector over *global* pressure dofs."""
        nl, no = self.n, self.part.n_owned
        p = v[2 * nl:]
        s = float(self._allreduce([float(p[:no].sum().item())])[0])
        p -= s / self.n_pressure_global

    def _test_nullspace(self):
        nl = self.n
        c = self.d_t
        c.zero_()
        c[2 * nl:] = 1.0 / math.sqrt(self.n_global)
        self.hemo.spmv(self.d_vals, c, self.d_w)
        self.hemo.mask_nodes(self.ghost_mask, self.d_w)
        r = self.gnorm(self.d_w)
        scale = math.sqrt(float(self._allreduce([self.hemo.dot(self.d_vals, self.d_vals)])[0])) / math.sqrt(3 * self.n_global)
        return bool(r < 1e-8 * max(scale, 1e-300))

    # ---- operators ---------------------------------------------------------------------------
    def _residual(self, x, out):
        """x must carry valid ghost values; out has zero ghost entries."""
        self.hemo.assemble_residual(x, self.d_un, self.d_bcval if self._has_bc else None, out)
        self.hemo.mask_nodes(self.ghost_mask, out)

    def _tic(self, key):
        """Optional wall-clock section timing (HEMO_DIST_PROFILE=1): synchronises the stream."""
        if not self._profile:
            return
        import time
        torch.cuda.synchronize(self.hemo.device)
        now = time.perf_counter()
        if self._last_key is not None:
            self.timers[self._last_key] = self.timers.get(self._last_key, 0.0) + now - self._last_t
        self._last_key, self._last_t = key, now

    def _fgmres(self, b, y):
        """Right-preconditioned FGMRES(restart) with global reductions; zero initial guess.
        b has zero ghosts; y gets valid ghost values."""
        hemo = self.hemo
        N, ldv, m = self.N, self.ldv, self.restart
        V, Z = self.V, self.Z
        y.zero_()
        bnorm = self.gnorm(b)
        if bnorm == 0.0:
            return 0, 0.0
        tol = max(self.ksp_rtol * bnorm, 1e-50)
        its = 0
        beta = bnorm
        hemo.vec_scale(1.0 / beta, b, V[:N])
        res = bnorm
        while its < self.ksp_max_it:
            H = np.zeros((m + 1, m))
            cs = np.zeros(m)
            sn = np.zeros(m)
            gvec = np.zeros(m + 1)
            gvec[0] = beta
            j = 0
            converged = False
            while j < m and its < self.ksp_max_it:
                vj = V[j * ldv:j * ldv + N]
                zj = Z[j * ldv:j * ldv + N]
                rin = vj
                if self.overlap > 1:
                    # restricted additive Schwarz: the local solve sees the residual on its overlap
                    self._tic("halo")
                    rin = self.d_t
                    rin.copy_(vj)
                    self.halo.update(rin)
                if self.global_pressure:
                    self._tic("schur_global")
                    self._global_schur(vj, zj)                    # z_p with global coupling (incl. ghosts)
                    self._tic("pc_u")
                    hemo.pc_apply(self.d_vals, rin, zj)           # z_u = A00_loc^-1 (r_u - A01 z_p)
                else:
                    hemo.pc_apply(self.d_vals, rin, zj)
                    if self._nullspace:
                        self._remove_pressure_mean(zj)
                if self.overlap > 1 or not self.global_pressure:
                    hemo.mask_nodes(self.ghost_mask, zj)          # keep the owned part only
                self._tic("halo")
                self.halo.update(zj)                              # ghost values from the owners
                self._tic("spmv")
                w = self.d_w
                hemo.spmv(self.d_vals, zj, w)
                hemo.mask_nodes(self.ghost_mask, w)
                self._tic("mdot+allreduce")
                # one reduction per iteration: w is stored as basis slot j+1, so the same multi-dot
                # returns V^T w and w.w; ||w - V h||^2 = w.w - |h|^2, recomputed exactly only when
                # cancellation makes it unreliable
                wslot = V[(j + 1) * ldv:(j + 1) * ldv + N]
                wslot.copy_(w)
                red = self._allreduce(hemo.vec_mdot(V, ldv, j + 2, w))
                h, ww = red[:j + 1], float(red[j + 1])
                self._tic("maxpy+allreduce")
                nsq_est = ww - float(h @ h)
                if self.ksp_rtol >= 1e-7 and nsq_est > 0.05 * ww:
                    hemo.vec_maxpy(V, ldv, h, -1.0, w)
                    hn = math.sqrt(nsq_est)
                else:
                    nsq = hemo.vec_maxpy(V, ldv, h, -1.0, w, want_normsq=True)
                    hn = math.sqrt(max(float(self._allreduce([nsq])[0]), 0.0))
                self._tic("host")
                H[:j + 1, j] = h
                H[j + 1, j] = hn
                if hn > 0.0:
                    hemo.vec_scale(1.0 / hn, w, V[(j + 1) * ldv:(j + 1) * ldv + N])
                for i in range(j):
                    a, b2 = H[i, j], H[i + 1, j]
                    H[i, j] = cs[i] * a + sn[i] * b2
                    H[i + 1, j] = -sn[i] * a + cs[i] * b2
                a, b2 = H[j, j], H[j + 1, j]
                d = math.hypot(a, b2)
                cs[j], sn[j] = (a / d, b2 / d) if d > 0 else (1.0, 0.0)
                H[j, j], H[j + 1, j] = d, 0.0
                gvec[j + 1] = -sn[j] * gvec[j]
                gvec[j] = cs[j] * gvec[j]
                its += 1
                res = abs(gvec[j + 1])
                if self.verbose and (its <= 5 or its % 10 == 0):
                    print(f"      KSP {its:4d} {res / bnorm:.3e}")
                j += 1
                if res <= tol or hn == 0.0:
                    converged = True
                    break
            k = j
            yk = np.linalg.solve(np.triu(H[:k, :k]), gvec[:k])
            hemo.vec_maxpy(Z, ldv, yk, 1.0, y)
            if converged:
                return its, res / bnorm
            # restart: r = b - A y (y has valid ghosts)
            hemo.spmv(self.d_vals, y, self.d_t)
            self.d_t.mul_(-1.0).add_(b)
            hemo.mask_nodes(self.ghost_mask, self.d_t)
            beta = self.gnorm(self.d_t)
            if beta <= tol:
                return its, beta / bnorm
            hemo.vec_scale(1.0 / beta, self.d_t, V[:N])
        raise RuntimeError("FGMRES reached max_it without converging")

    def _newton(self):
        hemo = self.hemo
        x, f, y, w, g = self.d_x, self.d_f, self.d_y, self.d_w, self.d_g
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
                kits, _ = self._fgmres(f, y)
            except RuntimeError:
                return it, lin_its, -3
            lin_its += kits
            hemo.spmv(self.d_vals, y, self.d_t)
            hemo.mask_nodes(self.ghost_mask, self.d_t)
            slope = self.gdot(f, self.d_t)
            slope = -abs(slope) if slope != 0.0 else -1.0
            alpha, lam = 1e-4, 1.0
            f2 = 0.5 * fnorm * fnorm
            lam_prev = g_prev = None
            accepted = False
            trial = torch.empty_like(x)
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
            w.copy_(y)
            hemo.mask_nodes(self.ghost_mask, w)
            ynorm = lam * self.gnorm(w)
            x.copy_(trial)
            f.copy_(g)
            fnorm = gnorm
            if self.verbose:
                print(f"  {it + 1} SNES Function norm {fnorm:.12e}")
            if fnorm < self.snes_atol:
                return it + 1, lin_its, 2
            if fnorm <= ttol:
                return it + 1, lin_its, 3
            w.copy_(x)
            hemo.mask_nodes(self.ghost_mask, w)
            if ynorm < self.snes_stol * self.gnorm(w):
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
        q = float(self._allreduce([q_loc])[0])
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
