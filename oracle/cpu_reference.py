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
