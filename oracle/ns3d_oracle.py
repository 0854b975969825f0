"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  **PARITY UNPINNED** (see ns_oracle.py).

Global 3-D restatement on P1–P1 tetrahedra built on `simplex_oracle`: assembly of the cell integrals
into the block system [u interleaved (3n) | p (n)], Dirichlet treatment of `assemble_matrix_block` /
`assemble_vector_block(..., x0=x, alpha=-1)` (src/solvers/stabilized_schur.py:144-175) and a Newton
iteration with sparse LU — what the 3-D CUDA path will be checked against once it has a solver, and
already used for a known-answer test on the Ethier–Steinman solution the reference's
src/scenarios/taylor_green.py:74-134 compares with (tests/test_ns3d_oracle.py).

Exterior-facet terms (stabilized_schur.py:79; stabilized_schur_pressure_backflow.py:189-217) enter
through `facet_sets` (ns_oracle.FacetSet objects, integrated by simplex_oracle.facet_F on triangular
facets with `facet_rule`).  On a mesh whose whole boundary carries Dirichlet velocity conditions the
term of :79 only touches constrained rows, so the Ethier-Steinman test runs without it.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import simplex_oracle as S

EPS0 = float(np.finfo(np.float64).resolution)


def unit_cube_tets(n: int):
    """n^3 cubes, 6 tetrahedra each (Kuhn subdivision): x (N, 3), cells (6 n^3, 4)."""
    g = np.linspace(0.0, 1.0, n + 1)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    x = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    base = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1)
    idx = lambda q: (q[:, 0] * (n + 1) + q[:, 1]) * (n + 1) + q[:, 2]
    cells = []
    for perm in ((0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)):
        v = [base.copy()]
        for ax in perm:
            w = v[-1].copy()
            w[:, ax] += 1
            v.append(w)
        cells.append(np.stack([idx(q) for q in v], axis=1))
    return x, np.concatenate(cells).astype(np.int32)


@dataclass
class Problem3D:
    x: np.ndarray                     # (n, 3)
    cells: np.ndarray                 # (E, 4)
    dt: float
    rho: float
    mu: float
    f: np.ndarray                     # (3,)
    rules: dict                       # 'Fu','Fp','uu','up','pu','pp' -> (pts, wts) on the reference tetrahedron
    bc_dofs: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))   # global dofs in [u (3n) | p (n)]
    facet_sets: list = field(default_factory=list)                                # ns_oracle.FacetSet
    facet_rule: tuple | None = None                                               # (pts (m,2), wts) on the reference triangle
    eps0: float = EPS0
    theta: float = 0.5
    a0: float = 1.0
    formulation: str = "standard"     # "curlcurl": stabilized_schur_pressurebc.py (oracle/curlcurl_oracle.py)

    def __post_init__(self):
        self.h = S.cell_diameter(self.x, self.cells)
        c = self.cells.astype(np.int64)
        n = self.n
        self.l2g = np.hstack([(3 * c[:, :, None] + np.arange(3)[None, None, :]).reshape(-1, 12), 3 * n + c])

    @property
    def n(self):
        return self.x.shape[0]

    @property
    def ndof(self):
        return 4 * self.n

    def _par(self):
        return dict(dt=self.dt, rho=self.rho, mu=self.mu, f=self.f, eps0=self.eps0, theta=self.theta, a0=self.a0)


def _kern(prob):
    if prob.formulation == "curlcurl":
        from .curlcurl_oracle import KernelsSimplex
        return KernelsSimplex
    return S


def assemble_F_raw(prob, xk, un, uh=None):
    S = _kern(prob)
    n = prob.n
    c = prob.cells
    U, P, Un = xk[:3 * n].reshape(-1, 3)[c], xk[3 * n:][c], un.reshape(-1, 3)[c]
    Uh = None if uh is None else uh.reshape(-1, 3)[c]
    Fu, _ = S.element_F(prob.x, c, prob.h, U, P, Un, prob.rules["Fu"], Uh=Uh, **prob._par())
    _, Fp = S.element_F(prob.x, c, prob.h, U, P, Un, prob.rules["Fp"], Uh=Uh, **prob._par())
    b = np.zeros(prob.ndof, dtype=Fu.dtype)
    np.add.at(b, prob.l2g[:, :12].reshape(-1), Fu.reshape(-1))
    np.add.at(b, prob.l2g[:, 12:].reshape(-1), Fp.reshape(-1))
    for fs in prob.facet_sets:
        ce = fs.pairs[:, 0]
        Ff = S.facet_F(prob.x, c, prob.h, fs.pairs, fs, U[ce], P[ce], Un[ce], prob.facet_rule, prob.rho, prob.mu,
                       prob.theta)
        np.add.at(b, prob.l2g[ce, :12].reshape(-1), Ff.reshape(-1))
    return b


def assemble_J_raw(prob, xk, un, uh=None):
    S = _kern(prob)
    n = prob.n
    c = prob.cells
    E = c.shape[0]
    U, P, Un = xk[:3 * n].reshape(-1, 3)[c], xk[3 * n:][c], un.reshape(-1, 3)[c]
    Uh = None if uh is None else uh.reshape(-1, 3)[c]
    kw = dict(Uh=Uh, **prob._par())
    Ae = np.zeros((E, 16, 16))
    Ae[:, :12, :12] = S.element_J(prob.x, c, prob.h, U, P, Un, prob.rules["uu"], **kw)[0].reshape(E, 12, 12)
    Ae[:, :12, 12:] = S.element_J(prob.x, c, prob.h, U, P, Un, prob.rules["up"], **kw)[1].reshape(E, 12, 4)
    Ae[:, 12:, :12] = S.element_J(prob.x, c, prob.h, U, P, Un, prob.rules["pu"], **kw)[2].reshape(E, 4, 12)
    Ae[:, 12:, 12:] = S.element_J(prob.x, c, prob.h, U, P, Un, prob.rules["pp"], **kw)[3]
    for fs in prob.facet_sets:
        ce = fs.pairs[:, 0]
        # (m, 12, 16): columns ordered (U flattened, then P) = the cell's local dof order
        np.add.at(Ae, (ce, slice(0, 12)), S.facet_J(prob.x, c, prob.h, fs.pairs, fs, Un[ce], prob.facet_rule, prob.rho,
                                                   prob.mu, prob.theta))
    rows = np.repeat(prob.l2g, 16, axis=1).reshape(-1)
    cols = np.tile(prob.l2g, (1, 16)).reshape(-1)
    A = sp.coo_matrix((Ae.reshape(-1), (rows, cols)), shape=(prob.ndof, prob.ndof)).tocsr()
    A.sort_indices()
    return A


def bc_arrays(prob, bc_lists=None):
    """marker (bool, ndof) and diagonal multiplicity (ndof): `bc_lists` = one dof array per Dirichlet
    condition (a dof held by k conditions gets k on the diagonal); default: the
    single list prob.bc_dofs."""
    lists = [prob.bc_dofs] if bc_lists is None else bc_lists
    mult = np.zeros(prob.ndof)
    for dofs in lists:
        mult[np.asarray(dofs, dtype=np.int64)] += 1.0
    return mult > 0, mult


def assemble_system(prob, xk, un, g, uh=None, bc_lists=None):
    """(A, b) of assemble_matrix_block / assemble_vector_block + apply_lifting(x0 = x, alpha = -1) +
    set_bc (stabilized_schur.py:144-175)."""
    marker, mult = bc_arrays(prob, bc_lists)
    A_raw = assemble_J_raw(prob, xk, un, uh)
    b = assemble_F_raw(prob, xk, un, uh)
    d = np.where(marker, g - xk, 0.0)
    b = b + A_raw @ d
    b[marker] = xk[marker] - g[marker]
    keep = sp.diags((~marker).astype(np.float64))
    A = (keep @ A_raw @ keep + sp.diags(mult)).tocsr()      # values only: explicit zeros of the FE pattern are dropped
    A.sort_indices()
    return A, b


def newton_step(prob, x0, un, g, uh=None, rtol=1e-10, max_it=20):
    """One time step: Newton with sparse LU on F(x) = 0 with x[bc] = g[bc] (Dirichlet rows/cols zeroed,
    unit diagonal, lifting with x0 = x, alpha = -1)."""
    x = x0.copy()
    marker = np.zeros(prob.ndof, dtype=bool)
    marker[prob.bc_dofs] = True
    keep = sp.diags((~marker).astype(np.float64))
    f0 = None
    for it in range(max_it):
        A_raw = assemble_J_raw(prob, x, un, uh)
        b = assemble_F_raw(prob, x, un, uh)
        d = np.where(marker, g - x, 0.0)
        b = b + A_raw @ d
        b[marker] = x[marker] - g[marker]
        fn = np.linalg.norm(b)
        if f0 is None:
            f0 = fn
        if fn <= rtol * max(f0, 1e-300) or fn < 1e-14:
            return x, it
        A = (keep @ A_raw @ keep + sp.diags(marker.astype(np.float64))).tocsc()
        x = x - spla.splu(A).solve(b)
    return x, max_it
