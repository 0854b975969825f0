"""CPU restatement of the reference's *solver configuration* (test / baseline infrastructure
only): SNES newtonls + bt, KSP FGMRES(200) rtol 1e-5, PC fieldsplit Schur FULL with SELFP,
sub-KSP u = GMRES(30, rtol 1e-5) + ILU, sub-KSP p = preonly + ILU
(/root/reference/src/solvers/stabilized_schur.py:202-275), on top of the C cell kernels
(oracle/c/p1tri_cells.c).  PETSc's ILU(0) is approximated by SuperLU's incomplete LU with
fill_factor 1 — this is a port for timing the algorithm on host cores, not PETSc itself
(parity unpinned; DOLFINx/PETSc cannot be installed here)."""
from __future__ import annotations

import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import ns_oracle as O
from .c_oracle import FastAssembler


def _fgmres(A, b, pc, rtol=1e-5, restart=200, maxit=1000):
    n = len(b)
    x = np.zeros(n)
    r = b.copy()
    beta = np.linalg.norm(r)
    b0 = beta
    its = 0
    if b0 == 0.0:
        return x, 0
    while its < maxit:
        m = min(restart, maxit - its)
        V = [r / beta]
        Z = []
        H = np.zeros((m + 1, m))
        g = np.zeros(m + 1)
        g[0] = beta
        cs, sn = [], []
        j = 0
        done = False
        while j < m:
            z = pc(V[j])
            Z.append(z)
            w = A @ z
            for i in range(j + 1):
                H[i, j] = V[i] @ w
            for i in range(j + 1):
                w = w - H[i, j] * V[i]
            H[j + 1, j] = np.linalg.norm(w)
            V.append(w / H[j + 1, j] if H[j + 1, j] > 0 else w)
            for i in range(j):
                t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
                H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
                H[i, j] = t
            d = np.hypot(H[j, j], H[j + 1, j])
            cs.append(H[j, j] / d)
            sn.append(H[j + 1, j] / d)
            H[j, j] = d
            H[j + 1, j] = 0.0
            g[j + 1] = -sn[j] * g[j]
            g[j] = cs[j] * g[j]
            its += 1
            j += 1
            if abs(g[j]) <= rtol * b0:
                done = True
                break
        y = np.linalg.solve(np.triu(H[:j, :j]), g[:j])
        for i in range(j):
            x += y[i] * Z[i]
        if done:
            return x, its
        r = b - A @ x
        beta = np.linalg.norm(r)
        if beta <= rtol * b0:
            return x, its
    raise RuntimeError("FGMRES did not converge")


class ReferenceLikeSolver:
    def __init__(self, prob: O.Problem):
        self.prob = prob
        self.asm = FastAssembler(prob)
        self.timers = {"assembly": 0.0, "pc_setup": 0.0, "ksp": 0.0}
        self.lin_its = 0

    def _linear_solve(self, A, f):
        n = self.prob.n
        t0 = time.perf_counter()
        A = A.tocsr()
        A00 = A[:2 * n, :2 * n].tocsc()
        A01 = A[:2 * n, 2 * n:].tocsr()
        A10 = A[2 * n:, :2 * n].tocsr()
        A11 = A[2 * n:, 2 * n:].tocsr()
        Sp = (A11 - A10 @ sp.diags(1.0 / A00.diagonal()) @ A01).tocsc()          # SELFP
        ilu0 = spla.spilu(A00, fill_factor=1.0, drop_tol=0.0)
        singular = O.has_constant_pressure_nullspace(self.prob, A)
        if singular:
            Sp = (Sp + 1e-10 * sp.identity(n) * abs(Sp.diagonal()).max()).tocsc()
        ilus = spla.spilu(Sp, fill_factor=1.0, drop_tol=0.0)
        M0 = spla.LinearOperator(A00.shape, ilu0.solve)
        self.timers["pc_setup"] += time.perf_counter() - t0

        def a00_inv(v):
            z, _ = spla.gmres(A00, v, M=M0, restart=30, rtol=1e-5, maxiter=20)
            return z

        def pc(r):                                # FULL Schur factorisation
            zu = a00_inv(r[:2 * n])
            rp = r[2 * n:] - A10 @ zu
            if singular:
                rp = rp - rp.mean()
            zp = ilus.solve(rp)
            if singular:
                zp = zp - zp.mean()
            zu = zu - a00_inv(A01 @ zp)
            return np.concatenate([zu, zp])

        t0 = time.perf_counter()
        y, its = _fgmres(A, f, pc)
        self.timers["ksp"] += time.perf_counter() - t0
        self.lin_its += its
        return y

    def newton(self, x0, un, rtol=1e-8, stol=1e-8, max_it=100):
        prob = self.prob
        n = prob.n
        x = x0.copy()
        t0 = time.perf_counter()
        f = self.asm.F(x, un)
        self.timers["assembly"] += time.perf_counter() - t0
        fnorm = np.linalg.norm(f)
        ttol = rtol * fnorm
        for it in range(max_it):
            t0 = time.perf_counter()
            A = self.asm.J(x[:2 * n], x[2 * n:], un)
            self.timers["assembly"] += time.perf_counter() - t0
            y = self._linear_solve(A, f)
            lam = 1.0
            for _ in range(10):                   # bt line search (first trial accepted in practice)
                w = x - lam * y
                t0 = time.perf_counter()
                g = self.asm.F(w, un)
                self.timers["assembly"] += time.perf_counter() - t0
                gnorm = np.linalg.norm(g)
                if gnorm <= fnorm or lam < 1e-3:
                    break
                lam *= 0.5
            ynorm = lam * np.linalg.norm(y)
            x, f, fnorm = w, g, gnorm
            if fnorm <= ttol or ynorm < stol * np.linalg.norm(x):
                return x, it + 1
        return x, max_it

    def step(self, x, un):
        x = O.remove_nullspace(self.prob, x)
        return self.newton(x, un)


# =====================================================================================================
# The same configuration on all host cores: C + OpenMP (oracle/c/ref_ksp.c), one thread per "MPI rank".
# This is what bench.py times as the CPU arm (`--impl reference`, `cpu_baseline`).
# =====================================================================================================
import ctypes as _C
import os as _os

from . import build_c as _build_c
from . import c_oracle as _CO

_ksp_lib = None


def _ksp():
    global _ksp_lib
    if _ksp_lib is None:
        L = _C.CDLL(_build_c.build(which="ksp"))
        L.refksp_create.restype = _C.c_void_p
        L.refksp_create.argtypes = [_C.c_int, _C.c_int, _C.c_int, _C.c_int]
        L.refksp_destroy.argtypes = [_C.c_void_p]
        L.refksp_set_inner.argtypes = [_C.c_void_p, _C.c_int, _C.c_int, _C.c_double]
        L.refksp_set_nullspace.argtypes = [_C.c_void_p, _C.c_int]
        L.refksp_stats.argtypes = [_C.c_void_p, _C.c_void_p]
        L.refksp_setup.argtypes = [_C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_void_p]
        L.refksp_solve.restype = _C.c_int
        L.refksp_solve.argtypes = [_C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_double, _C.c_int, _C.c_void_p, _C.c_void_p]
        L.ref_insert_matrix.argtypes = [_C.c_int64, _C.c_int, _C.c_void_p, _C.c_void_p, _C.c_int64, _C.c_void_p, _C.c_int]
        L.ref_insert_vector.argtypes = [_C.c_int64, _C.c_int, _C.c_void_p, _C.c_void_p, _C.c_int64, _C.c_void_p, _C.c_int]
        L.ref_cell_positions.argtypes = [_C.c_int64, _C.c_int, _C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_int]
        _ksp_lib = L
    return _ksp_lib


def _ptr(a):
    return a.ctypes.data_as(_C.c_void_p)


def block_pattern(nrowptr, ncol):
    """CSR pattern create_matrix_block builds for [u interleaved (2n) | p (n)] from the sorted node graph
    (stabilized_schur.py:191): every row of node i holds (2j, 2j+1) for its neighbours j, then 2n + j."""
    nrowptr = np.asarray(nrowptr, dtype=np.int64)
    ncol = np.asarray(ncol, dtype=np.int64)
    n = nrowptr.shape[0] - 1
    deg = np.diff(nrowptr)
    nnz_node = int(nrowptr[-1])
    rowlen = np.empty(3 * n, dtype=np.int64)
    rowlen[0:2 * n:2] = 3 * deg
    rowlen[1:2 * n:2] = 3 * deg
    rowlen[2 * n:] = 3 * deg
    rowptr = np.zeros(3 * n + 1, dtype=np.int64)
    np.cumsum(rowlen, out=rowptr[1:])
    colind = np.empty(9 * nnz_node, dtype=np.int32)
    rows = np.repeat(np.arange(n, dtype=np.int64), deg)
    t = np.arange(nnz_node, dtype=np.int64) - nrowptr[rows]
    d = deg[rows]
    for base in (6 * nrowptr[rows], 6 * nrowptr[rows] + 3 * d, 6 * nnz_node + 3 * nrowptr[rows]):
        colind[base + 2 * t] = 2 * ncol
        colind[base + 2 * t + 1] = 2 * ncol + 1
        colind[base + 2 * d + t] = 2 * n + ncol
    return rowptr, colind


class CReferenceSolver:
    """SNES newtonls + bt (oracle/ns_oracle.newton_solve) around the C restatement of
    KSP fgmres(200) / fieldsplit Schur FULL + SELFP / gmres(30)+asm-ilu0 / preonly+asm-ilu0
    (/root/reference/src/solvers/stabilized_schur.py:202-275).  P1 triangles.  `nranks` threads stand for the
    MPI ranks of `mpirun -n nranks` (default: all cores).  `sub_pc="lu"` is the configuration of the hemodynamic variants
    (stabilized_schur_pressure_backflow.py:255-297: gmres/lu on A00, preonly/lu on Sp): SciPy's SuperLU, sequential like
    PETSc's own `lu` (a parallel PETSc run needs an external package), factorised for every Jacobian; the ASM/ILU(0) blocks
    do not converge on the stenosis workloads (1000 outer iterations)."""

    def __init__(self, prob: O.Problem, nranks: int | None = None, nullspace: bool | None = None,
                 ksp_rtol: float = 1e-5, ksp_max_it: int = 1000, restart: int = 200, node_graph=None, sub_pc: str = "ilu"):
        from cfd_hemodynamic_b200.fem import discretization as D
        L = _ksp()
        self.prob = prob
        self.nranks = int(nranks or (_os.cpu_count() or 1))
        n = prob.n
        self.n = n
        if prob.cells.shape[1] != 3:
            raise NotImplementedError("the CPU arm is written for P1 triangles")
        nrowptr, ncol = node_graph if node_graph is not None else D.node_graph(prob.cells, n)
        self.rowptr, self.colind = block_pattern(nrowptr, ncol)
        self.nnz = int(self.rowptr[-1])
        self.vals = np.zeros(self.nnz)
        self.l2g = np.ascontiguousarray(O.local_to_global(prob), dtype=np.int64)
        E = prob.cells.shape[0]
        self.pos = np.empty((E, 9, 9), dtype=np.int32)
        L.ref_cell_positions(E, 9, _ptr(self.l2g), _ptr(self.rowptr), _ptr(self.colind), _ptr(self.pos), self.nranks)
        assert (self.pos >= 0).all()
        self.marker, self.g, self.mult = O.bc_arrays(prob)
        if self.marker.any():
            md = np.nonzero(self.marker)[0]
            rows = np.concatenate([np.arange(self.rowptr[r], self.rowptr[r + 1]) for r in md])
            cols = np.nonzero(self.marker[self.colind])[0]
            self.zero_idx = np.unique(np.concatenate([rows, cols]))
            self.diag_idx = np.array([self.rowptr[r] + np.searchsorted(self.colind[self.rowptr[r]:self.rowptr[r + 1]], r)
                                      for r in md], dtype=np.int64)
            self.diag_val = self.mult[md]
        self.ksp = L.refksp_create(2 * n, n, self.nranks, restart)
        self.ksp_rtol, self.ksp_max_it = ksp_rtol, ksp_max_it
        # "ilu": gmres(30)+asm/ilu(0) and preonly+asm/ilu(0) in C (stabilized_schur.py:256-267);
        # "lu":  gmres+lu and preonly+lu (stabilized_schur_pressure_backflow.py:284-288) with SciPy's SuperLU — sequential
        #        like PETSc's own `lu` (a parallel run needs MUMPS / SuperLU_DIST); factorised for every Jacobian.
        if sub_pc not in ("ilu", "lu"):
            raise ValueError(sub_pc)
        self.sub_pc = sub_pc
        self.restart = restart
        self._lu_outer = 0
        self._nullspace = nullspace
        self.timers = {"assembly": 0.0, "pc_setup": 0.0, "ksp": 0.0}
        self.lin_its = 0
        self.newton_its = 0

    def close(self):
        if self.ksp:
            _ksp().refksp_destroy(self.ksp)
            self.ksp = None

    # --- assembly: C cell kernels on all cores, threaded insertion into the fixed pattern ---------
    def J_raw_vals(self, u, p, un, out):
        prob = self.prob
        Ae, _ = _CO.element_tensors(prob, u, p, un, True, False)
        _ksp().ref_insert_matrix(Ae.shape[0], 9, _ptr(Ae), _ptr(self.pos), self.nnz, _ptr(out), self.nranks)
        for fs in prob.facet_sets:
            ce = fs.pairs[:, 0]
            Af = O.facet_matrices(prob, fs, un)                       # (m, 6, 9), boundary-sized
            np.add.at(out, self.pos[ce][:, :6, :].reshape(-1), Af.reshape(-1))
        return out

    def J(self, u, p, un):
        t0 = time.perf_counter()
        self.J_raw_vals(u, p, un, self.vals)
        if self.marker.any():
            self.vals[self.zero_idx] = 0.0
            self.vals[self.diag_idx] = self.diag_val
        self.timers["assembly"] += time.perf_counter() - t0
        A = sp.csr_matrix((self.vals, self.colind, self.rowptr), shape=(3 * self.n, 3 * self.n), copy=False)
        return A

    def F(self, x, un):
        t0 = time.perf_counter()
        prob = self.prob
        n = self.n
        u, p = x[:2 * n], x[2 * n:]
        _, Fe = _CO.element_tensors(prob, u, p, un, False, True)
        b = np.empty(3 * n)
        _ksp().ref_insert_vector(Fe.shape[0], 9, _ptr(Fe), _ptr(self.l2g), 3 * n, _ptr(b), self.nranks)
        if prob.facet_sets:
            U, P, Un = O._gather(prob, u, p, un)
            for fs in prob.facet_sets:
                ce = fs.pairs[:, 0]
                Fu_f = O.facet_F(prob, fs, U[ce], P[ce], Un[ce])
                np.add.at(b, self.l2g[ce][:, :6].reshape(-1), Fu_f.reshape(-1))
        if self.marker.any():
            d = np.where(self.marker, self.g - x, 0.0)
            if np.any(d != 0.0):
                tmp = np.zeros(self.nnz)
                self.J_raw_vals(u, p, un, tmp)
                b = b + sp.csr_matrix((tmp, self.colind, self.rowptr), shape=(3 * n, 3 * n)) @ d
            b[self.marker] = x[self.marker] - self.g[self.marker]
        self.timers["assembly"] += time.perf_counter() - t0
        return b

    # --- KSPSolve -----------------------------------------------------------------------------------
    def _linear_solve_lu(self, A, f):
        """FGMRES(restart) + fieldsplit Schur FULL, SELFP, exact (LU) sub-solves: the inner gmres on A00 converges in one
        iteration with an exact preconditioner, so it is applied directly."""
        n = self.n
        t0 = time.perf_counter()
        A = A.tocsr()
        A00 = A[:2 * n, :2 * n].tocsc()
        A01 = A[:2 * n, 2 * n:].tocsr()
        A10 = A[2 * n:, :2 * n].tocsr()
        A11 = A[2 * n:, 2 * n:].tocsr()
        Sp = (A11 - A10 @ sp.diags(1.0 / A00.diagonal()) @ A01).tocsc()          # SELFP (:253)
        singular = bool(self._nullspace)
        if singular:
            Sp = (Sp + 1e-10 * abs(Sp.diagonal()).max() * sp.identity(n)).tocsc()
        lu0 = spla.splu(A00)
        lus = spla.splu(Sp)
        t1 = time.perf_counter()

        def pc(r):
            zu = lu0.solve(r[:2 * n])
            rp = r[2 * n:] - A10 @ zu
            if singular:
                rp = rp - rp.mean()
            zp = lus.solve(rp)
            if singular:
                zp = zp - zp.mean()
            zu = zu - lu0.solve(A01 @ zp)
            return np.concatenate([zu, zp])

        y, its = _fgmres(A, np.asarray(f, dtype=np.float64), pc, rtol=self.ksp_rtol, restart=self.restart, maxit=self.ksp_max_it)
        self.timers["pc_setup"] += t1 - t0
        self.timers["ksp"] += time.perf_counter() - t1
        self.lin_its += its
        self._lu_outer += its
        return y

    def linear_solve(self, A, f):
        L = _ksp()
        if self._nullspace is None:
            self._nullspace = bool(O.has_constant_pressure_nullspace(self.prob, A))
        if self.sub_pc == "lu":
            return self._linear_solve_lu(A, f)
        L.refksp_set_nullspace(self.ksp, int(self._nullspace))
        t0 = time.perf_counter()
        L.refksp_setup(self.ksp, _ptr(self.rowptr), _ptr(self.colind), _ptr(self.vals))
        t1 = time.perf_counter()
        y = np.empty(3 * self.n)
        its = _C.c_int(0)
        rel = _C.c_double(0.0)
        f = np.ascontiguousarray(f)
        rc = L.refksp_solve(self.ksp, _ptr(f), _ptr(y), self.ksp_rtol, self.ksp_max_it, _C.byref(its), _C.byref(rel))
        self.timers["pc_setup"] += t1 - t0
        self.timers["ksp"] += time.perf_counter() - t1
        self.lin_its += its.value
        if rc != 0:
            raise RuntimeError(f"reference KSP did not converge in {its.value} iterations (rel. residual {rel.value:.2e})")
        return y

    def stats(self):
        if self.sub_pc == "lu":
            return dict(outer_its=int(self._lu_outer), inner_its=int(self._lu_outer), inner_solves=int(2 * self._lu_outer))
        out = np.zeros(3, dtype=np.int64)
        _ksp().refksp_stats(self.ksp, _ptr(out))
        return dict(outer_its=int(out[0]), inner_its=int(out[1]), inner_solves=int(out[2]))

    def step(self, x, un, rtol=1e-8, stol=1e-8):
        """solveStep (stabilized_schur.py:313-334): nullsp.remove(x_n) then SNES.solve."""
        x = O.remove_nullspace(self.prob, x)
        x, its, reason = O.newton_solve(self.prob, x, un, rtol=rtol, stol=stol, asm=self, linear_solve=self.linear_solve)
        self.newton_its += its
        if reason < 0:
            raise RuntimeError(f"Did not converge, reason: {reason}.")
        return x
