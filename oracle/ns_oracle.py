"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  **PARITY UNPINNED.**

A plain numpy/scipy fp64 restatement of the reference's per-timestep hot path
for P1–P1 triangles.  It exists to check the CUDA path; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it.  The product (`cfd_hemodynamic_b200`) never does.

"Parity unpinned": the arithmetic of the reference lives in DOLFINx v0.9 /
FFCx / Basix / PETSc (pinned by reference singularity.def:2), none of which is
installable here, and the reference ships no tests, golden vectors or fixtures
for this path.  The oracle is therefore anchored on the
reference's own call sites (cited per function) plus known-answer checks
(Jacobian == derivative of the residual by complex step, patch tests,
Poiseuille, constant-pressure null space), not on reference outputs.

What is restated (file:line under /root/reference):
  * residual F           src/solvers/stabilized_schur.py:60-123
  * J = derivative(F)    src/solvers/stabilized_schur.py:185-189
  * per-block quadrature `form(extract_blocks(...))` :188-189 — each block is
    its own UFL form, so its own estimated degree: F_u, J_uu 12; F_p, J_up,
    J_pu 11; J_pp 10 (P1 triangles)
  * BC semantics         :144-175 (assemble_matrix_block / assemble_vector_block
    with x0=x, alpha=-1: zeroed rows/cols, +1 diagonal per DirichletBC, lifting,
    set_bc)
  * boundary-facet terms src/solvers/stabilized_schur_pressure_backflow.py:170-217,
    src/solvers/stabilized_schur_backflow.py:158-176
  * Newton + bt search   PETSc SNES newtonls as configured at :202-213,269-275
  * sigma/epsilon        src/solverBase.py:176-182

Element integrals are evaluated FFCx-style: the full integrand at every
quadrature point (no moment factorisation), vectorised over cells.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

EPS0 = float(np.finfo(np.float64).resolution)  # 1e-15, stabilized_schur.py:100

# reference P1 gradients on the reference triangle {1-x-y, x, y}
_GHAT = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])


@dataclass
class FacetSet:
    """One tagged exterior-facet integral.  Term coefficients (all default 0):

      a_p     * p (n.v)                          stabilized_schur.py:79
      pconst  * (n.v)                            pressure_backflow.py:193,208
      a_g     * -mu ((nabla_grad u_m) n).v       stabilized_schur.py:79
      a_s     * -(2 mu eps(u_m) n).v             pressure_backflow.py:209
      a_n     * [-(2mu eps(u_m) n).v_T - (2mu eps(v) n).u_T
                 + (beta_n mu/h) u_T.v_T]        pressure_backflow.py:195-201
      a_b     * -beta_b rho (u_n.n)_- (u_m.v)    pressure_backflow.py:214-217
    The a_* carry the multiplicity of the term (setup() called twice adds the
    boundary terms twice).
    """
    pairs: np.ndarray                 # (m, 2) int32 (cell, local facet)
    a_p: float = 0.0
    pconst: float = 0.0
    a_g: float = 0.0
    a_s: float = 0.0
    a_n: float = 0.0
    beta_n: float = 0.0
    a_b: float = 0.0
    beta_b: float = 0.0


@dataclass
class Problem:
    x: np.ndarray                     # (n, 2)
    cells: np.ndarray                 # (E, 3) int32 triangles | (E, 4) tensor-ordered quadrilaterals
    h: np.ndarray                     # (E,)
    dt: float
    rho: float
    mu: float
    f: np.ndarray                     # (2,)
    rules: dict                       # 'Fu','Fp','uu','up','pu','pp' -> (pts, wts)
    facet_rule: tuple                 # (pts on [0,1], wts)
    facet_sets: list = field(default_factory=list)
    # Dirichlet: list of (block 'u'|'p', unrolled dofs within block, g block vector)
    bcs: list = field(default_factory=list)
    eps0: float = EPS0
    # Time scheme.  The forms are evaluated at u_e = theta u + (1 - theta) u_n with the time
    # derivative (a0 u - u_h) / dt.  Default = the mid-point scheme of stabilized_schur.py:69-80
    # (theta = 1/2, a0 = 1, u_h = u_n); stabilized_schur_bdf2.py:76-110 uses theta = 1 and
    # a0 u + a1 u_n + a2 u_nn, i.e. u_h = -(a1 u_n + a2 u_nn) (nodal vector `uh`, 2n).
    # tau_supg / tau_lsic always take u_n (u_prev).
    theta: float = 0.5
    a0: float = 1.0
    uh: np.ndarray | None = None
    formulation: str = "standard"     # "curlcurl": stabilized_schur_pressurebc.py (oracle/curlcurl_oracle.py)

    @property
    def n(self):
        return self.x.shape[0]

    @property
    def ndof(self):
        return 3 * self.n


# --------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------
def cell_geometry(x, cells):
    X = x[cells]                                   # (E,3,2)
    J = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]], axis=2)   # J[:, i, j] = dx_i/dxi_j
    det = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
    inv = np.empty_like(J)
    inv[:, 0, 0] = J[:, 1, 1] / det
    inv[:, 0, 1] = -J[:, 0, 1] / det
    inv[:, 1, 0] = -J[:, 1, 0] / det
    inv[:, 1, 1] = J[:, 0, 0] / det
    # grad phi_a = J^{-T} ghat_a  → dphi[e, a, i] = sum_j ghat[a, j] inv[e, j, i]
    dphi = np.einsum("aj,eji->eai", _GHAT, inv)
    return np.abs(det), dphi


def cell_diameter(x, cells):
    """mesh.h for triangles: longest edge (stabilized_schur.py:83-88)."""
    X = x[cells]
    e = [np.linalg.norm(X[:, i] - X[:, j], axis=1) for i, j in ((0, 1), (0, 2), (1, 2))]
    return np.maximum(np.maximum(e[0], e[1]), e[2])


def _tau(prob, unq, h):
    """tau_supg and tau_lsic at one quadrature point (stabilized_schur.py:91-118)."""
    nu = prob.mu / prob.rho
    vnorm = np.sqrt(np.einsum("ei,ei->e", unq, unq))
    two_v = 2.0 * vnorm
    t1 = h / np.where(two_v >= prob.eps0, two_v, prob.eps0)
    t2 = prob.dt / 2.0
    t3 = (h * h) / (4.0 * nu)
    tau = (1.0 / t1 ** 2 + 1.0 / t2 ** 2 + 1.0 / t3 ** 2) ** (-0.5)
    Re = (vnorm * h) / (2.0 * nu)
    z = np.where(Re <= 3.0, Re / 3.0, 1.0)
    tau_l = (vnorm * h * z) / 2.0
    return tau, tau_l


# --------------------------------------------------------------------------
# cell integrals
# --------------------------------------------------------------------------
def element_F(prob, U, P, Un, rule, Uh=None):
    """Element residual (Fu (E,3,2), Fp (E,3)) with the given cell rule.
    U, Un: (E,3,2) nodal values per cell (may be complex for the complex-step
    check), P: (E,3); Uh: history values of the time derivative (default Un)."""
    det, dphi = cell_geometry(prob.x, prob.cells)
    pts, wts = rule
    rho, mu, dt = prob.rho, prob.mu, prob.dt
    th, a0 = prob.theta, prob.a0
    Uh = Un if Uh is None else Uh
    f = np.asarray(prob.f, dtype=np.float64)
    Um = th * U + (1.0 - th) * Un
    G = np.einsum("eai,eaj->eij", dphi, Um)          # G_ij = d_i u_mj (nabla_grad)
    gradp = np.einsum("eai,ea->ei", dphi, P)
    divu = G[:, 0, 0] + G[:, 1, 1]
    eps = 0.5 * (G + np.swapaxes(G, 1, 2))
    Fu = np.zeros(U.shape, dtype=U.dtype)
    Fp = np.zeros(P.shape, dtype=U.dtype)
    for q in range(len(wts)):
        xi, eta = pts[q]
        phi = np.array([1.0 - xi - eta, xi, eta])
        w = wts[q] * det                               # (E,)
        u = np.einsum("a,eai->ei", phi, U)
        un = np.einsum("a,eai->ei", phi, Un)
        um = th * u + (1.0 - th) * un
        p = np.einsum("a,ea->e", phi, P)
        tau, tau_l = _tau(prob, un.real, prob.h)
        conv = np.einsum("ei,eij->ej", um, G)           # (u_m . nabla) u_m
        dudt = (a0 * u - np.einsum("a,eai->ei", phi, Uh)) / dt
        sigma = 2.0 * mu * eps - p[:, None, None] * np.eye(2)[None]
        R = rho * (dudt + conv) + gradp - rho * f[None, :]     # -div sigma = grad p (P1)
        um_dphi = np.einsum("ei,eai->ea", um, dphi)     # u_m . grad phi_a
        # :74-77 Galerkin momentum
        Fu += w[:, None, None] * (
            rho * phi[None, :, None] * (dudt + conv - f[None, :])[:, None, :]
            + np.einsum("eai,eik->eak", dphi, sigma))
        # :109 SUPG, :119 LSIC
        Fu += w[:, None, None] * (
            tau[:, None, None] * um_dphi[:, :, None] * R[:, None, :]
            + (tau_l * rho * divu)[:, None, None] * dphi)
        # :80 continuity, :113 PSPG
        Fp += w[:, None] * (phi[None, :] * divu[:, None]
                            + (tau / rho)[:, None] * np.einsum("ei,eai->ea", R, dphi))
    return Fu, Fp


def element_J(prob, U, P, Un, rule, Uh=None):
    """Element Jacobian blocks by hand-differentiated integrand:
    Juu (E,3,2,3,2) [a,k ; b,l], Jup (E,3,2,3), Jpu (E,3,3,2), Jpp (E,3,3).
    With the general time scheme d(u_e)/du = theta and d(dudt)/du = a0/dt (theta = 1/2,
    a0 = 1 reproduce the factors 1/2 and 1/dt of the mid-point scheme)."""
    det, dphi = cell_geometry(prob.x, prob.cells)
    pts, wts = rule
    rho, mu, dt = prob.rho, prob.mu, prob.dt
    th, a0 = prob.theta, prob.a0
    Uh = Un if Uh is None else Uh
    f = np.asarray(prob.f, dtype=np.float64)
    E = prob.cells.shape[0]
    Um = th * U + (1.0 - th) * Un
    G = np.einsum("eai,eaj->eij", dphi, Um)
    gradp = np.einsum("eai,ea->ei", dphi, P)
    I2 = np.eye(2)
    Juu = np.zeros((E, 3, 2, 3, 2))
    Jup = np.zeros((E, 3, 2, 3))
    Jpu = np.zeros((E, 3, 3, 2))
    Jpp = np.zeros((E, 3, 3))
    # viscous: eps(v):2mu eps(du/2) = mu eps(v):eps(du)
    #   eps(v):eps(du) for v=phi_a e_k, du=phi_b e_l = 1/2 (dphi_a.dphi_b d_kl + d_l phi_a d_k phi_b)
    dd = np.einsum("eai,ebi->eab", dphi, dphi)
    visc = th * mu * (dd[:, :, None, :, None] * I2[None, None, :, None, :]
                      + np.einsum("eal,ebk->eakbl", dphi, dphi))
    for q in range(len(wts)):
        xi, eta = pts[q]
        phi = np.array([1.0 - xi - eta, xi, eta])
        w = wts[q] * det
        u = np.einsum("a,eai->ei", phi, U)
        un = np.einsum("a,eai->ei", phi, Un)
        um = th * u + (1.0 - th) * un
        tau, tau_l = _tau(prob, un, prob.h)
        conv = np.einsum("ei,eij->ej", um, G)
        R = rho * ((a0 * u - np.einsum("a,eai->ei", phi, Uh)) / dt + conv) + gradp - rho * f[None, :]
        um_dphi = np.einsum("ei,eai->ea", um, dphi)
        # dR_k/d(u_b,l) = rho [ a0 phi_b/dt d_kl + theta phi_b G_lk + theta (um.dphi_b) d_kl ]
        dR = rho * ((a0 * phi[None, :] / dt + th * um_dphi)[:, None, :, None] * I2[None, :, None, :]
                    + th * phi[None, None, :, None] * np.swapaxes(G, 1, 2)[:, :, None, :])   # (E,k,b,l)
        # Galerkin: rho phi_a [phi_b/dt d_kl + 1/2 phi_b G_lk + 1/2 um.dphi_b d_kl] = phi_a dR
        Juu_q = np.einsum("a,ekbl->eakbl", phi, dR) + visc
        # SUPG: tau [ dR_k (um.dphi_a) + theta R_k phi_b d_l phi_a ]
        Juu_q = Juu_q + tau[:, None, None, None, None] * (
            np.einsum("ea,ekbl->eakbl", um_dphi, dR)
            + th * np.einsum("ek,b,eal->eakbl", R, phi, dphi))
        # LSIC: theta tau_l rho d_l phi_b d_k phi_a
        Juu_q = Juu_q + (th * tau_l * rho)[:, None, None, None, None] * np.einsum(
            "eak,ebl->eakbl", dphi, dphi)
        Juu += w[:, None, None, None, None] * Juu_q
        # J_up: -phi_b d_k phi_a  + tau d_k phi_b (um.dphi_a)
        Jup += w[:, None, None, None] * (
            -np.einsum("b,eak->eakb", phi, dphi)
            + tau[:, None, None, None] * np.einsum("ea,ebk->eakb", um_dphi, dphi))
        # J_pu: theta phi_a d_l phi_b + (tau/rho) dR_k,bl d_k phi_a
        Jpu += w[:, None, None, None] * (
            th * np.einsum("a,ebl->eabl", phi, dphi)
            + (tau / rho)[:, None, None, None] * np.einsum("ekbl,eak->eabl", dR, dphi))
        # J_pp: (tau/rho) dphi_a.dphi_b
        Jpp += (w * tau / rho)[:, None, None] * dd
    return Juu, Jup, Jpu, Jpp


# --------------------------------------------------------------------------
# exterior-facet integrals
# --------------------------------------------------------------------------
_FACET_VERTS = np.array([[1, 2], [0, 2], [0, 1]])   # facet i opposite vertex i


def facet_F(prob, fs: FacetSet, U, P, Un):
    """Element residual contributions (Fu (m,3,2)) of one facet set; linear in
    (U, P) apart from the constant-pressure term."""
    cells = prob.cells[fs.pairs[:, 0]]
    lf = fs.pairs[:, 1]
    X = prob.x[cells]                                   # (m,3,2)
    _, dphi = cell_geometry(prob.x, cells)               # boundary cells only
    h = prob.h[fs.pairs[:, 0]]
    m = cells.shape[0]
    ar = np.arange(m)
    va = _FACET_VERTS[lf, 0]
    vb = _FACET_VERTS[lf, 1]
    xa, xb = X[ar, va], X[ar, vb]
    xo = X[ar, lf]                                      # opposite vertex
    t = xb - xa
    length = np.linalg.norm(t, axis=1)
    nrm = np.stack([t[:, 1], -t[:, 0]], axis=1) / length[:, None]
    # outward: pointing away from the opposite vertex
    sgn = np.sign(np.einsum("ei,ei->e", nrm, xa - xo))
    nrm = nrm * sgn[:, None]
    mu, rho = prob.mu, prob.rho
    Um = prob.theta * U + (1.0 - prob.theta) * Un
    G = np.einsum("eai,eaj->eij", dphi, Um)             # d_i u_mj
    eps = 0.5 * (G + np.swapaxes(G, 1, 2))
    Gn = np.einsum("eij,ej->ei", G, nrm)                # (nabla_grad u) n
    en = np.einsum("eij,ej->ei", eps, nrm)              # eps(u) n
    Fu = np.zeros(U.shape, dtype=U.dtype)
    pts, wts = prob.facet_rule
    Pn = np.eye(2)[None] - nrm[:, :, None] * nrm[:, None, :]     # tangential projector
    # eps(v) n for v = phi_a e_k: 1/2 (d_i phi_a n_k + (dphi_a.n) d_ik)  → (m,a,k,i)
    dn = np.einsum("eai,ei->ea", dphi, nrm)
    epsv_n = 0.5 * (np.einsum("eai,ek->eaki", dphi, nrm) + dn[:, :, None, None] * np.eye(2)[None, None])
    for q in range(len(wts)):
        s = pts[q]
        w = wts[q] * length
        phi = np.zeros((m, 3))
        phi[ar, va] = 1.0 - s
        phi[ar, vb] = s
        um = np.einsum("ea,eai->ei", phi, Um)
        un = np.einsum("ea,eai->ei", phi, Un).real
        p = np.einsum("ea,ea->e", phi, P)
        umT = np.einsum("eij,ej->ei", Pn, um)
        # v = phi_a e_k ; v_T = phi_a Pn[:,k]
        val = np.zeros(U.shape, dtype=U.dtype)
        val += (fs.a_p * p + fs.pconst)[:, None, None] * phi[:, :, None] * nrm[:, None, :]
        val -= fs.a_g * mu * phi[:, :, None] * Gn[:, None, :]
        val -= fs.a_s * 2.0 * mu * phi[:, :, None] * en[:, None, :]
        if fs.a_n != 0.0:
            enT = np.einsum("eik,ei->ek", Pn, en)                       # (2mu eps n).v_T → project
            val -= fs.a_n * 2.0 * mu * phi[:, :, None] * enT[:, None, :]
            val -= fs.a_n * 2.0 * mu * np.einsum("eaki,ei->eak", epsv_n, umT)
            val += fs.a_n * (fs.beta_n * mu / h)[:, None, None] * phi[:, :, None] * umT[:, None, :]
        if fs.a_b != 0.0:
            unn = np.einsum("ei,ei->e", un, nrm)
            un_minus = 0.5 * (unn - np.abs(unn))
            val -= fs.a_b * fs.beta_b * rho * un_minus[:, None, None] * phi[:, :, None] * um[:, None, :]
        Fu += w[:, None, None] * val
    return Fu


def outlet_flux(prob, pairs, Un_nodal):
    """Q = int u_prev . n ds over the given facets
    (pressure_backflow.py:204-211, 383-385)."""
    if prob.cells.shape[1] == 4:
        from . import q1_oracle
        return q1_oracle.outlet_flux(prob, pairs, Un_nodal)
    if prob.cells.shape[1] == 6:
        from . import pk_oracle
        return pk_oracle.outlet_flux(prob, pairs, Un_nodal)
    cells = prob.cells[pairs[:, 0]]
    lf = pairs[:, 1]
    X = prob.x[cells]
    ar = np.arange(cells.shape[0])
    va, vb = _FACET_VERTS[lf, 0], _FACET_VERTS[lf, 1]
    xa, xb, xo = X[ar, va], X[ar, vb], X[ar, lf]
    t = xb - xa
    nrm = np.stack([t[:, 1], -t[:, 0]], axis=1)          # |nrm| = length
    sgn = np.sign(np.einsum("ei,ei->e", nrm, xa - xo))
    nrm = nrm * sgn[:, None]
    Uc = Un_nodal.reshape(-1, 2)[cells]
    umid = 0.5 * (Uc[ar, va] + Uc[ar, vb])
    return float(np.sum(np.einsum("ei,ei->e", umid, nrm)))


# --------------------------------------------------------------------------
# global assembly
# --------------------------------------------------------------------------
def _gather(prob, u, p, un):
    c = prob.cells
    return u.reshape(-1, 2)[c], p[c], un.reshape(-1, 2)[c]


def _gather_history(prob):
    """Cell values of the history vector of the time derivative (None: u_n is used)."""
    return None if prob.uh is None else np.asarray(prob.uh).reshape(-1, 2)[prob.cells]


def _kernels(prob):
    """Element routines for the problem's cell type: this module for P1
    triangles, oracle/q1_oracle.py for Q1 quadrilaterals (4 nodes per cell); `prob.formulation = "curlcurl"` selects
    the rotational form of stabilized_schur_pressurebc.py (oracle/curlcurl_oracle.py, P1 triangles)."""
    if getattr(prob, "formulation", "standard") == "curlcurl":
        assert prob.cells.shape[1] == 3
        from .curlcurl_oracle import Kernels2D
        return Kernels2D
    if prob.cells.shape[1] == 4:
        from . import q1_oracle
        return q1_oracle
    if prob.cells.shape[1] == 6:          # P2-P2 triangles (3 vertex + 3 edge nodes per cell)
        from . import pk_oracle
        return pk_oracle
    import sys
    return sys.modules[__name__]


def local_to_global(prob):
    """(E, 3*nv) global dof of each element-local dof; local order
    [u(a=0,k=0), u(0,1), u(1,0), ..., p0, p1, ...]; global [u interleaved | p]."""
    c = prob.cells.astype(np.int64)
    n = prob.n
    nv = c.shape[1]
    ud = (2 * c[:, :, None] + np.arange(2)[None, None, :]).reshape(-1, 2 * nv)
    pd = 2 * n + c
    return np.hstack([ud, pd])


def sparsity_pattern(prob):
    """Full FE pattern, sorted unique columns per row (3P create_matrix_block;
    trigger stabilized_schur.py:191)."""
    l2g = local_to_global(prob)
    nl = l2g.shape[1]
    rows = np.repeat(l2g, nl, axis=1).reshape(-1)
    cols = np.tile(l2g, (1, nl)).reshape(-1)
    A = sp.coo_matrix((np.ones(rows.shape[0]), (rows, cols)), shape=(prob.ndof, prob.ndof)).tocsr()
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int32)


def assemble_F_raw(prob, u, p, un):
    """Residual without Dirichlet treatment (cells + facets)."""
    K = _kernels(prob)
    U, P, Un = _gather(prob, u, p, un)
    nv = prob.cells.shape[1]
    Uh = _gather_history(prob)
    Fu, _ = K.element_F(prob, U, P, Un, prob.rules["Fu"], Uh)
    _, Fp = K.element_F(prob, U, P, Un, prob.rules["Fp"], Uh)
    b = np.zeros(prob.ndof, dtype=Fu.dtype)
    l2g = local_to_global(prob)
    np.add.at(b, l2g[:, :2 * nv].reshape(-1), Fu.reshape(-1))
    np.add.at(b, l2g[:, 2 * nv:].reshape(-1), Fp.reshape(-1))
    for fs in prob.facet_sets:
        ce = fs.pairs[:, 0]
        Fu_f = K.facet_F(prob, fs, U[ce], P[ce], Un[ce])
        np.add.at(b, l2g[ce][:, :2 * nv].reshape(-1), Fu_f.reshape(-1))
    return b


def element_matrices(prob, u, p, un):
    """(E, 3nv, 3nv) cell element matrices with per-block quadrature."""
    K = _kernels(prob)
    U, P, Un = _gather(prob, u, p, un)
    E, nv = prob.cells.shape
    nu = 2 * nv
    Ae = np.zeros((E, 3 * nv, 3 * nv))
    Uh = _gather_history(prob)
    Juu, _, _, _ = K.element_J(prob, U, P, Un, prob.rules["uu"], Uh)
    _, Jup, _, _ = K.element_J(prob, U, P, Un, prob.rules["up"], Uh)
    _, _, Jpu, _ = K.element_J(prob, U, P, Un, prob.rules["pu"], Uh)
    _, _, _, Jpp = K.element_J(prob, U, P, Un, prob.rules["pp"], Uh)
    Ae[:, :nu, :nu] = Juu.reshape(E, nu, nu)
    Ae[:, :nu, nu:] = Jup.reshape(E, nu, nv)
    Ae[:, nu:, :nu] = Jpu.reshape(E, nv, nu)
    Ae[:, nu:, nu:] = Jpp
    return Ae


def facet_matrices(prob, fs: FacetSet, un):
    """(m, 2nv, 3nv) d(Fu_facet)/d(U,P): the facet terms are affine in (U,P), so
    the exact derivative is obtained column by column from unit vectors."""
    K = _kernels(prob)
    ce = fs.pairs[:, 0]
    m = ce.shape[0]
    nv = prob.cells.shape[1]
    nu = 2 * nv
    Unc = un.reshape(-1, 2)[prob.cells[ce]]
    Z2 = np.zeros((m, nv, 2))
    Z1 = np.zeros((m, nv))
    F0 = K.facet_F(prob, fs, Z2, Z1, Unc)
    out = np.zeros((m, nu, 3 * nv))
    for j in range(3 * nv):
        U = Z2.copy()
        P = Z1.copy()
        if j < nu:
            U[:, j // 2, j % 2] = 1.0
        else:
            P[:, j - nu] = 1.0
        out[:, :, j] = (K.facet_F(prob, fs, U, P, Unc) - F0).reshape(m, nu)
    return out


def assemble_J_raw(prob, u, p, un):
    """Jacobian (CSR, full FE pattern, explicit zeros kept) without BCs."""
    Ae = element_matrices(prob, u, p, un)
    l2g = local_to_global(prob)
    nl = l2g.shape[1]
    nu = 2 * prob.cells.shape[1]
    rows = np.repeat(l2g, nl, axis=1).reshape(-1)
    cols = np.tile(l2g, (1, nl)).reshape(-1)
    vals = Ae.reshape(-1)
    for fs in prob.facet_sets:
        ce = fs.pairs[:, 0]
        Af = facet_matrices(prob, fs, un)
        lg = l2g[ce]
        rows = np.concatenate([rows, np.repeat(lg[:, :nu], nl, axis=1).reshape(-1)])
        cols = np.concatenate([cols, np.tile(lg, (1, nu)).reshape(-1)])
        vals = np.concatenate([vals, Af.reshape(-1)])
    A = sp.coo_matrix((vals, (rows, cols)), shape=(prob.ndof, prob.ndof)).tocsr()
    A.sort_indices()
    return A


# --------------------------------------------------------------------------
# Dirichlet treatment (3P semantics of assemble_*_block)
# --------------------------------------------------------------------------
def bc_arrays(prob):
    """marker (bool, ndof), g (ndof; last BC in the list wins), diagonal
    multiplicity (number of DirichletBC objects containing the dof)."""
    n = prob.n
    marker = np.zeros(prob.ndof, dtype=bool)
    g = np.zeros(prob.ndof)
    mult = np.zeros(prob.ndof)
    for block, dofs, gvec in prob.bcs:
        off = 0 if block == "u" else 2 * n
        d = np.asarray(dofs, dtype=np.int64) + off
        marker[d] = True
        g[d] = np.asarray(gvec)[np.asarray(dofs, dtype=np.int64)]
        mult[d] += 1.0
    return marker, g, mult


def assemble_J(prob, u, p, un):
    """assemble_matrix_block(J, J_form, bcs) (stabilized_schur.py:144-155)."""
    A = assemble_J_raw(prob, u, p, un)
    marker, _, mult = bc_arrays(prob)
    if marker.any():
        keep = sp.diags((~marker).astype(np.float64))
        A = (keep @ A @ keep).tocsr() + sp.diags(mult)
        A = A.tocsr()
        A.sort_indices()
    return A


def assemble_F(prob, x, un):
    """assemble_vector_block(F, F_form, J_form, bcs, x0=x, alpha=-1)
    (stabilized_schur.py:157-175): b = F(x) + A_raw[:, bc] (g - x)[bc];
    b[bc] = x[bc] - g[bc]."""
    n = prob.n
    u, p = x[:2 * n], x[2 * n:]
    b = assemble_F_raw(prob, u, p, un)
    marker, g, _ = bc_arrays(prob)
    if marker.any():
        d = np.where(marker, g - x, 0.0)
        if np.any(d != 0.0):
            A = assemble_J_raw(prob, u, p, un)
            b = b + A @ d
        b[marker] = x[marker] - g[marker]
    return b


# --------------------------------------------------------------------------
# Newton with cubic backtracking (PETSc SNES newtonls + bt)
# --------------------------------------------------------------------------
def remove_nullspace(prob, x):
    """nullsp.remove(x): subtract the mean over pressure dofs
    (stabilized_schur.py:283-293,319)."""
    n = prob.n
    x = x.copy()
    x[2 * n:] -= x[2 * n:].mean()
    return x


def has_constant_pressure_nullspace(prob, A, tol=1e-8):
    """nullsp.test(A) (stabilized_schur.py:314)."""
    n = prob.n
    v = np.zeros(prob.ndof)
    v[2 * n:] = 1.0 / np.sqrt(n)
    return np.linalg.norm(A @ v) < tol * max(1.0, spla.norm(A, np.inf))


def newton_solve(prob, x0, un, rtol=1e-8, atol=1e-50, stol=1e-8, max_it=100,
                 verbose=False, asm=None, linear_solve=None):
    """One SNES.solve (stabilized_schur.py:321) with an exact (sparse LU)
    linear solve.  Returns (x, iterations, reason>0 converged).  `asm`: optional
    assembler object with J(u, p, un) / F(x, un) (oracle/c_oracle.FastAssembler).
    `linear_solve(A, f) -> y`: optional replacement of the sparse-LU solve (oracle/cpu_reference.py
    plugs in the restated FGMRES + fieldsplit configuration of the reference)."""
    n = prob.n
    x = x0.copy()
    _F = (lambda xx: asm.F(xx, un)) if asm is not None else (lambda xx: assemble_F(prob, xx, un))
    _J = (lambda xx: asm.J(xx[:2 * n], xx[2 * n:], un)) if asm is not None else \
        (lambda xx: assemble_J(prob, xx[:2 * n], xx[2 * n:], un))
    f = _F(x)
    fnorm = np.linalg.norm(f)
    ttol = rtol * fnorm
    if verbose:
        print(f"  0 SNES Function norm {fnorm:.12e}")
    if fnorm < atol:
        return x, 0, 2
    for it in range(max_it):
        A = _J(x)
        singular = linear_solve is None and has_constant_pressure_nullspace(prob, A)
        if linear_solve is not None:
            y = linear_solve(A, f)
        elif singular:
            # pin through a bordered system: solve in the orthogonal complement
            e = np.zeros(prob.ndof); e[2 * n:] = 1.0 / np.sqrt(n)
            B = sp.bmat([[A, sp.csr_matrix(e[:, None])], [sp.csr_matrix(e[None, :]), None]]).tocsc()
            y = spla.splu(B).solve(np.concatenate([f, [0.0]]))[:-1]
        else:
            y = spla.splu(A.tocsc()).solve(f)
        # bt line search
        slope = -abs(float(f @ (A @ y)))
        if slope == 0.0:
            slope = -1.0
        lam = 1.0
        alpha = 1e-4
        accepted = False
        lam_prev = g_prev = None
        f2 = 0.5 * fnorm ** 2
        for _ in range(40):
            w = x - lam * y
            g = _F(w)
            gnorm = np.linalg.norm(g)
            if 0.5 * gnorm ** 2 <= f2 + lam * alpha * slope:
                accepted = True
                break
            g2 = 0.5 * gnorm ** 2
            if lam_prev is None:
                lam_new = -slope / (2.0 * (g2 - f2 - slope)) if lam == 1.0 else lam * 0.5
                lam_new = min(max(lam_new, 0.1 * lam), 0.5 * lam)
            else:
                t1 = g2 - f2 - lam * slope
                t2 = g_prev - f2 - lam_prev * slope
                a = (t1 / lam ** 2 - t2 / lam_prev ** 2) / (lam - lam_prev)
                b = (-lam_prev * t1 / lam ** 2 + lam * t2 / lam_prev ** 2) / (lam - lam_prev)
                d = b * b - 3 * a * slope
                d = max(d, 0.0)
                lam_new = -slope / (2 * b) if a == 0 else (-b + np.sqrt(d)) / (3 * a)
                lam_new = min(max(lam_new, 0.1 * lam), 0.5 * lam)
            lam_prev, g_prev = lam, g2
            lam = lam_new
        if not accepted:
            return x, it + 1, -6
        ynorm = lam * np.linalg.norm(y)
        x, f, fnorm = w, g, gnorm
        if verbose:
            print(f"  {it + 1} SNES Function norm {fnorm:.12e}")
        if fnorm < atol:
            return x, it + 1, 2
        if fnorm <= ttol:
            return x, it + 1, 3
        if ynorm < stol * np.linalg.norm(x):
            return x, it + 1, 4
    return x, max_it, -5
