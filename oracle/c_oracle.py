"""ctypes wrapper of the C restatement of the cell kernels (oracle/c/p1tri_cells.c).
TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/ns_oracle.py header: parity unpinned)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

from . import build_c
from . import ns_oracle as O

_lib = None
_RULE_ORDER = ("Fu", "Fp", "uu", "up", "pu", "pp")


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_c.build())
        _lib.hemo_ref_cells.restype = None
    return _lib


def element_tensors(prob: O.Problem, u, p, un, want_J=True, want_F=True):
    """(Ae (E,9,9) | None, Fe (E,9) | None) over all cells, OpenMP-parallel."""
    lib = _load()
    E = prob.cells.shape[0]
    x = np.ascontiguousarray(prob.x, dtype=np.float64)
    cells = np.ascontiguousarray(prob.cells, dtype=np.int32)
    h = np.ascontiguousarray(prob.h, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    un = np.ascontiguousarray(un, dtype=np.float64)
    f = np.ascontiguousarray(prob.f, dtype=np.float64)
    pts = [np.ascontiguousarray(prob.rules[k][0], dtype=np.float64) for k in _RULE_ORDER]
    wts = [np.ascontiguousarray(prob.rules[k][1], dtype=np.float64) for k in _RULE_ORDER]
    nq = np.array([len(w) for w in wts], dtype=np.int32)
    PP = C.POINTER(C.c_double)
    pts_arr = (PP * 6)(*[a.ctypes.data_as(PP) for a in pts])
    wts_arr = (PP * 6)(*[a.ctypes.data_as(PP) for a in wts])
    Ae = np.empty((E, 9, 9)) if want_J else None
    Fe = np.empty((E, 9)) if want_F else None
    vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    lib.hemo_ref_cells(C.c_int(E), vp(x), vp(cells), vp(h), vp(u), vp(p), vp(un), C.c_double(prob.dt),
                       C.c_double(prob.rho), C.c_double(prob.mu), vp(f), C.c_double(prob.eps0), vp(nq),
                       pts_arr, wts_arr, vp(Ae), vp(Fe))
    return Ae, Fe


class FastAssembler:
    """assemble_J / assemble_F with the C cell kernels and precomputed COO→CSR plumbing;
    facets, lifting and Dirichlet treatment reuse the numpy oracle (boundary-sized)."""

    def __init__(self, prob: O.Problem):
        self.prob = prob
        self.l2g = O.local_to_global(prob)
        self.rows = np.repeat(self.l2g, 9, axis=1).reshape(-1)
        self.cols = np.tile(self.l2g, (1, 9)).reshape(-1)
        self.marker, self.g, self.mult = O.bc_arrays(prob)
        self.keep = sp.diags((~self.marker).astype(np.float64))

    def J_raw(self, u, p, un):
        prob = self.prob
        Ae, _ = element_tensors(prob, u, p, un, True, False)
        rows, cols, vals = self.rows, self.cols, Ae.reshape(-1)
        for fs in prob.facet_sets:
            ce = fs.pairs[:, 0]
            Af = O.facet_matrices(prob, fs, un)
            lg = self.l2g[ce]
            rows = np.concatenate([rows, np.repeat(lg[:, :6], 9, axis=1).reshape(-1)])
            cols = np.concatenate([cols, np.tile(lg, (1, 6)).reshape(-1)])
            vals = np.concatenate([vals, Af.reshape(-1)])
        A = sp.coo_matrix((vals, (rows, cols)), shape=(prob.ndof, prob.ndof)).tocsr()
        return A

    def J(self, u, p, un):
        A = self.J_raw(u, p, un)
        if self.marker.any():
            A = (self.keep @ A @ self.keep + sp.diags(self.mult)).tocsr()
        return A

    def F(self, x, un):
        prob = self.prob
        n = prob.n
        u, p = x[:2 * n], x[2 * n:]
        _, Fe = element_tensors(prob, u, p, un, False, True)
        b = np.bincount(self.l2g.reshape(-1), weights=Fe.reshape(-1), minlength=prob.ndof)
        if prob.facet_sets:
            U, P, Un = O._gather(prob, u, p, un)
            for fs in prob.facet_sets:
                ce = fs.pairs[:, 0]
                Fu_f = O.facet_F(prob, fs, U[ce], P[ce], Un[ce])
                np.add.at(b, self.l2g[ce][:, :6].reshape(-1), Fu_f.reshape(-1))
        if self.marker.any():
            d = np.where(self.marker, self.g - x, 0.0)
            if np.any(d != 0.0):
                b = b + self.J_raw(u, p, un) @ d
            b[self.marker] = x[self.marker] - self.g[self.marker]
        return b
