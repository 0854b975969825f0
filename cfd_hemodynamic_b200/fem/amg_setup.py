"""Host-side, one-time construction of the multigrid transfer operators.

No reference equivalent: the reference preconditions A00 and Sp with PETSc
ASM/ILU(0) (src/solvers/stabilized_schur.py:256-267).  North-star asks for an
aggregation-AMG V-cycle instead.  The prolongators depend only on the mesh
(scalar P1 Laplacian), so they are built once here with scipy (symbolic work,
like DOLFINx's pattern builder) and handed to the library together with the
fixed patterns of A*P and R*A*P; the numeric Galerkin products run on the GPU
every Newton iteration.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from .._lib import load_library


def _aggregate(S: sp.csr_matrix, exclude: np.ndarray):
    lib = load_library()
    n = S.shape[0]
    rp = np.ascontiguousarray(S.indptr, dtype=np.int32)
    ci = np.ascontiguousarray(S.indices, dtype=np.int32)
    ex = np.ascontiguousarray(exclude, dtype=np.uint8)
    agg = np.empty(n, dtype=np.int32)
    na = C.c_int(0)
    rc = lib.hemo_host_aggregate(n, rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p),
                                 ex.ctypes.data_as(C.c_void_p), agg.ctypes.data_as(C.c_void_p), C.byref(na))
    if rc != 0:
        raise RuntimeError(f"hemo_host_aggregate failed ({rc})")
    return agg, na.value


def _strength(A: sp.csr_matrix, theta: float) -> sp.csr_matrix:
    A = A.tocoo()
    d = np.abs(A.tocsr().diagonal())
    d[d == 0.0] = 1.0
    keep = (A.row != A.col) & (np.abs(A.data) > theta * np.sqrt(d[A.row] * d[A.col]))
    S = sp.csr_matrix((np.ones(int(keep.sum())), (A.row[keep], A.col[keep])), shape=A.shape)
    S.sort_indices()
    return S


def build_hierarchy(L: sp.csr_matrix, dirichlet: np.ndarray, *, max_coarse: int, theta: float = 0.08,
                    smooth: bool = True, max_levels: int = 14, fine_pattern: sp.csr_matrix | None = None):
    """L: scalar P1 Laplacian on the node graph (fine pattern == node graph).
    dirichlet: bool mask of constrained nodes (their rows of P are empty).
    Returns a list of dicts with P, R (csr), AP and C patterns (csr of ones),
    following the level chain n_0 → n_1 → ... until n_l <= max_coarse."""
    n = L.shape[0]
    if fine_pattern is None:
        fine_pattern = sp.csr_matrix((np.ones(L.nnz), L.indices, L.indptr), shape=L.shape)
    else:       # level-0 operator lives on a wider pattern than the aggregation graph (SELFP)
        fine_pattern = sp.csr_matrix((np.ones(fine_pattern.nnz), fine_pattern.indices, fine_pattern.indptr),
                                     shape=fine_pattern.shape)
    free = (~dirichlet).astype(np.float64)
    K = sp.diags(free) @ L @ sp.diags(free)            # Dirichlet rows/cols removed
    K = (K + sp.diags(dirichlet.astype(np.float64) * L.diagonal())).tocsr()
    levels = []
    A = K
    Apat = fine_pattern
    excl = dirichlet.copy()
    for lev in range(max_levels):
        nl = A.shape[0]
        if nl <= max_coarse:
            break
        # dense coarse operators lose "strong" couplings: relax the threshold before giving up
        for th in (theta, 0.25 * theta, 0.0):
            S = _strength(A, th)
            # vertices without any strong coupling are left to the smoother (no coarse-grid
            # correction, empty row of P) instead of becoming singleton aggregates
            lonely = np.diff(S.indptr) == 0
            agg, na = _aggregate(S, excl | lonely)
            if 0 < na < 0.75 * nl:
                break
        else:
            if na == 0:
                raise RuntimeError(f"AMG coarsening found no couplings at level {lev} ({nl} vertices)")
            raise RuntimeError(f"AMG coarsening stalled at level {lev}: {nl} -> {na}")
        rows = np.nonzero(agg >= 0)[0]
        T = sp.csr_matrix((np.ones(rows.shape[0]), (rows, agg[rows])), shape=(nl, na))
        if smooth:
            d = A.diagonal().copy()
            d[d == 0.0] = 1.0
            Dinv = sp.diags(1.0 / d)
            rho = float(np.max(np.abs(Dinv @ A).sum(axis=1)))      # Gershgorin bound
            P = (T - (4.0 / (3.0 * rho)) * (Dinv @ (A @ T))).tocsr()
            if excl.any():
                P = (sp.diags((~excl).astype(np.float64)) @ P).tocsr()
            P.eliminate_zeros()
        else:
            P = T
        P.sort_indices()
        R = P.T.tocsr()
        R.sort_indices()
        Ppat = sp.csr_matrix((np.ones(P.nnz), P.indices, P.indptr), shape=P.shape)
        APpat = (Apat @ Ppat).tocsr()
        APpat.sort_indices()
        Cpat = (Ppat.T.tocsr() @ APpat).tocsr()
        Cpat.sort_indices()
        APpat.data[:] = 1.0
        Cpat.data[:] = 1.0
        levels.append(dict(P=P, R=R, AP=APpat, C=Cpat))
        A = (R @ A @ P).tocsr()
        A.sort_indices()
        Apat = Cpat
        excl = np.zeros(na, dtype=bool)
    else:
        raise RuntimeError("AMG hierarchy needs more than max_levels levels")
    if A.shape[0] > max_coarse:
        raise RuntimeError("AMG coarsest level too large")
    return levels
