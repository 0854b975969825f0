"""CPU ORACLE for Pk-Pk triangles, k = 1, 2 — TEST INFRASTRUCTURE ONLY (see oracle/ns_oracle.py header: parity unpinned
against DOLFINx/FFCx, which cannot be installed here).

Restates the same forms as oracle/ns_oracle.py for Lagrange degree `p_grade` in both spaces — what the reference
builds with `initVelocitySpace("Lagrange", cell, p_grade, ...)` / `initPressureSpace("Lagrange", cell, p_grade)` in
    /root/reference/src/solvers/stabilized_schur_pressure_backflow.py:71,102-161   (cell integrals)
    /root/reference/src/solvers/stabilized_schur_pressure_backflow.py:170-217      (boundary terms)
    /root/reference/src/solvers/stabilized_schur_backflow.py:63,85-176
FFCx-style: the full integrand at every quadrature point, nothing factorised.  For k >= 2 the strong residual keeps
the viscous part of -div sigma(u_m, p) = -mu (Laplace u_m + grad div u_m) + grad p, whose second
derivatives are constant on an affine P2 cell.

Local dof order (3P, Basix): vertices 0, 1, 2, then one dof per edge, edge i opposite vertex i: (1,2), (0,2), (0,1).
`prob.cells` is (E, 6): the three vertex nodes then the three edge nodes; `prob.x` holds the coordinates of all nodes
(vertices first is NOT assumed: the geometry is read from the first three nodes of every cell).  For k = 1 the
routines reproduce oracle/ns_oracle.py to rounding (tests/test_pk_oracle.py), which anchors the generalisation.
"""
from __future__ import annotations

import numpy as np

from . import ns_oracle as O

_FACET_VERTS = np.array([[1, 2], [0, 2], [0, 1]])
_VERT_REF = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])


def nodes_per_cell(k: int) -> int:
    return {1: 3, 2: 6}[k]


def degree_of(cells) -> int:
    return {3: 1, 6: 2}[cells.shape[1]]


def tabulate(k: int, pts):
    """phi (nq, nn), reference gradient (nq, nn, 2), reference Hessian (nq, nn, 2, 2) at reference points."""
    pts = np.asarray(pts, dtype=float).reshape(-1, 2)
    xi, eta = pts[:, 0], pts[:, 1]
    l = np.stack([1.0 - xi - eta, xi, eta], axis=1)                     # barycentric
    dl = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])               # d lambda_a / d (xi, eta)
    nq = pts.shape[0]
    if k == 1:
        return l, np.broadcast_to(dl, (nq, 3, 2)).copy(), np.zeros((nq, 3, 2, 2))
    phi = np.empty((nq, 6))
    dphi = np.empty((nq, 6, 2))
    hess = np.zeros((nq, 6, 2, 2))
    for a in range(3):
        phi[:, a] = l[:, a] * (2.0 * l[:, a] - 1.0)
        dphi[:, a] = (4.0 * l[:, a] - 1.0)[:, None] * dl[a][None, :]
        hess[:, a] = 4.0 * np.outer(dl[a], dl[a])[None]
    for e, (i, j) in enumerate(_FACET_VERTS):
        phi[:, 3 + e] = 4.0 * l[:, i] * l[:, j]
        dphi[:, 3 + e] = 4.0 * (l[:, i, None] * dl[j][None, :] + l[:, j, None] * dl[i][None, :])
        hess[:, 3 + e] = 4.0 * (np.outer(dl[i], dl[j]) + np.outer(dl[j], dl[i]))[None]
    return phi, dphi, hess


def cell_geometry(x, cells):
    """|det J| (E,) and J^-1 (E,2,2) with inv[e, j, i] = d xi_j / d x_i, from the three vertex nodes."""
    X = x[cells[:, :3]]
    J = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]], axis=2)
    det = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
    inv = np.empty_like(J)
    inv[:, 0, 0] = J[:, 1, 1] / det
    inv[:, 0, 1] = -J[:, 0, 1] / det
    inv[:, 1, 0] = -J[:, 1, 0] / det
    inv[:, 1, 1] = J[:, 0, 0] / det
    return np.abs(det), inv


def cell_diameter(x, cells):
    return O.cell_diameter(x, cells[:, :3])


def _point_fields(prob, k, inv, tab, q, U, P, Un, Uh):
    """Everything the integrands need at quadrature point q (vectorised over cells)."""
    phi, dref, href = tab
    th, a0, dt, rho, mu = prob.theta, prob.a0, prob.dt, prob.rho, prob.mu
    f = np.asarray(prob.f, dtype=np.float64)
    ph = phi[q]                                                     # (nn,)
    dphi = np.einsum("aj,eji->eai", dref[q], inv)                   # physical gradients (E, nn, 2)
    hphi = np.einsum("ajm,eji,eml->eail", href[q], inv, inv)        # physical Hessians (E, nn, 2, 2)
    Um = th * U + (1.0 - th) * Un
    u = np.einsum("a,eai->ei", ph, U)
    un = np.einsum("a,eai->ei", ph, Un)
    um = th * u + (1.0 - th) * un
    p = np.einsum("a,ea->e", ph, P)
    G = np.einsum("eai,eaj->eij", dphi, Um)                         # G_ij = d_i u_mj
    gradp = np.einsum("eai,ea->ei", dphi, P)
    divu = G[:, 0, 0] + G[:, 1, 1]
    eps = 0.5 * (G + np.swapaxes(G, 1, 2))
    lap = np.einsum("eaii,eaj->ej", hphi, Um)                       # Laplace u_m
    graddiv = np.einsum("eaij,eaj->ei", hphi, Um)                   # grad (div u_m)
    tau, tau_l = O._tau(prob, un.real, prob.h)
    conv = np.einsum("ei,eij->ej", um, G)
    dudt = (a0 * u - np.einsum("a,eai->ei", ph, Uh)) / dt
    R = rho * (dudt + conv) - mu * (lap + graddiv) + gradp - rho * f[None, :]
    um_dphi = np.einsum("ei,eai->ea", um, dphi)
    return dict(phi=ph, dphi=dphi, hphi=hphi, u=u, un=un, um=um, p=p, G=G, gradp=gradp, divu=divu, eps=eps, tau=tau,
                tau_l=tau_l, conv=conv, dudt=dudt, R=R, um_dphi=um_dphi)


def element_F(prob, U, P, Un, rule, Uh=None):
    """Element residual (Fu (E,nn,2), Fp (E,nn)); U, Un: (E,nn,2), P: (E,nn)."""
    k = degree_of(prob.cells)
    det, inv = cell_geometry(prob.x, prob.cells)
    pts, wts = rule
    tab = tabulate(k, pts)
    rho, mu = prob.rho, prob.mu
    f = np.asarray(prob.f, dtype=np.float64)
    Uh = Un if Uh is None else Uh
    Fu = np.zeros(U.shape, dtype=U.dtype)
    Fp = np.zeros(P.shape, dtype=U.dtype)
    for q in range(len(wts)):
        w = wts[q] * det
        c = _point_fields(prob, k, inv, tab, q, U, P, Un, Uh)
        phi, dphi = c["phi"], c["dphi"]
        sigma = 2.0 * mu * c["eps"] - c["p"][:, None, None] * np.eye(2)[None]
        Fu += w[:, None, None] * (
            rho * phi[None, :, None] * (c["dudt"] + c["conv"] - f[None, :])[:, None, :]
            + np.einsum("eai,eik->eak", dphi, sigma)
            + c["tau"][:, None, None] * c["um_dphi"][:, :, None] * c["R"][:, None, :]
            + (c["tau_l"] * rho * c["divu"])[:, None, None] * dphi)
        Fp += w[:, None] * (phi[None, :] * c["divu"][:, None]
                            + (c["tau"] / rho)[:, None] * np.einsum("ei,eai->ea", c["R"], dphi))
    return Fu, Fp


def element_J(prob, U, P, Un, rule, Uh=None):
    """Hand-differentiated element Jacobian blocks Juu (E,nn,2,nn,2) [a,k ; b,l], Jup (E,nn,2,nn), Jpu (E,nn,nn,2),
    Jpp (E,nn,nn); checked against the complex-step derivative of element_F (tests/test_pk_oracle.py)."""
    k = degree_of(prob.cells)
    det, inv = cell_geometry(prob.x, prob.cells)
    pts, wts = rule
    tab = tabulate(k, pts)
    rho, mu, dt = prob.rho, prob.mu, prob.dt
    th, a0 = prob.theta, prob.a0
    Uh = Un if Uh is None else Uh
    E, nn = prob.cells.shape
    I2 = np.eye(2)
    Juu = np.zeros((E, nn, 2, nn, 2))
    Jup = np.zeros((E, nn, 2, nn))
    Jpu = np.zeros((E, nn, nn, 2))
    Jpp = np.zeros((E, nn, nn))
    for q in range(len(wts)):
        w = wts[q] * det
        c = _point_fields(prob, k, inv, tab, q, U, P, Un, Uh)
        phi, dphi, hphi, G, tau, tau_l = c["phi"], c["dphi"], c["hphi"], c["G"], c["tau"], c["tau_l"]
        dd = np.einsum("eai,ebi->eab", dphi, dphi)
        visc = th * mu * (dd[:, :, None, :, None] * I2[None, None, :, None, :] + np.einsum("eal,ebk->eakbl", dphi, dphi))
        # dR_k/d(u_b,l) = rho [a0 phi_b/dt d_kl + theta phi_b G_lk + theta (um.dphi_b) d_kl]
        #               - theta mu [Laplace(phi_b) d_kl + d_k d_l phi_b]
        lapb = hphi[:, :, 0, 0] + hphi[:, :, 1, 1]
        dR = (rho * ((a0 * phi[None, :] / dt + th * c["um_dphi"])[:, None, :, None] * I2[None, :, None, :]
                     + th * phi[None, None, :, None] * np.swapaxes(G, 1, 2)[:, :, None, :])
              - th * mu * (lapb[:, None, :, None] * I2[None, :, None, :] + np.einsum("ebkl->ekbl", hphi)))
        # momentum Galerkin: only the rho part of dR multiplies phi_a (the viscous Galerkin term is `visc`)
        dRg = rho * ((a0 * phi[None, :] / dt + th * c["um_dphi"])[:, None, :, None] * I2[None, :, None, :]
                     + th * phi[None, None, :, None] * np.swapaxes(G, 1, 2)[:, :, None, :])
        Juu_q = np.einsum("a,ekbl->eakbl", phi, dRg) + visc
        Juu_q = Juu_q + tau[:, None, None, None, None] * (
            np.einsum("ea,ekbl->eakbl", c["um_dphi"], dR) + th * np.einsum("ek,b,eal->eakbl", c["R"], phi, dphi))
        Juu_q = Juu_q + (th * tau_l * rho)[:, None, None, None, None] * np.einsum("eak,ebl->eakbl", dphi, dphi)
        Juu += w[:, None, None, None, None] * Juu_q
        Jup += w[:, None, None, None] * (
            -np.einsum("b,eak->eakb", phi, dphi) + tau[:, None, None, None] * np.einsum("ea,ebk->eakb", c["um_dphi"], dphi))
        Jpu += w[:, None, None, None] * (
            th * np.einsum("a,ebl->eabl", phi, dphi) + (tau / rho)[:, None, None, None] * np.einsum("ekbl,eak->eabl", dR, dphi))
        Jpp += (w * tau / rho)[:, None, None] * dd
    return Juu, Jup, Jpu, Jpp


# ---------------------------------------------------------------------------------------------------------------------
# exterior-facet integrals: the cell basis and its gradients evaluated on the facet (all nn functions have non-zero
# gradients there for k = 2)
# ---------------------------------------------------------------------------------------------------------------------
def facet_F(prob, fs, U, P, Un):
    """Element residual contributions (m, nn, 2) of one facet set (same terms and coefficients as
    oracle/ns_oracle.facet_F)."""
    k = degree_of(prob.cells)
    cells = prob.cells[fs.pairs[:, 0]]
    lf = fs.pairs[:, 1]
    X = prob.x[cells[:, :3]]
    _, inv = cell_geometry(prob.x, cells)
    h = prob.h[fs.pairs[:, 0]]
    m = cells.shape[0]
    ar = np.arange(m)
    va, vb = _FACET_VERTS[lf, 0], _FACET_VERTS[lf, 1]
    xa, xb, xo = X[ar, va], X[ar, vb], X[ar, lf]
    t = xb - xa
    length = np.linalg.norm(t, axis=1)
    nrm = np.stack([t[:, 1], -t[:, 0]], axis=1) / length[:, None]
    nrm = nrm * np.sign(np.einsum("ei,ei->e", nrm, xa - xo))[:, None]
    mu, rho = prob.mu, prob.rho
    Um = prob.theta * U + (1.0 - prob.theta) * Un
    Pn = np.eye(2)[None] - nrm[:, :, None] * nrm[:, None, :]
    pts, wts = prob.facet_rule
    Fu = np.zeros(U.shape, dtype=U.dtype)
    for q in range(len(wts)):
        s = pts[q]
        ref = (1.0 - s) * _VERT_REF[va] + s * _VERT_REF[vb]          # (m, 2) reference point on the facet
        phi = np.empty((m, cells.shape[1]))
        dphi = np.empty((m, cells.shape[1], 2))
        for lfv in range(3):                                          # tabulate per local facet (3 distinct points)
            sel = lf == lfv
            if not sel.any():
                continue
            ph, dr, _ = tabulate(k, ref[sel][:1])
            phi[sel] = ph[0][None, :]
            dphi[sel] = np.einsum("aj,eji->eai", dr[0], inv[sel])
        w = wts[q] * length
        um = np.einsum("ea,eai->ei", phi, Um)
        un = np.einsum("ea,eai->ei", phi, Un).real
        p = np.einsum("ea,ea->e", phi, P)
        G = np.einsum("eai,eaj->eij", dphi, Um)
        eps = 0.5 * (G + np.swapaxes(G, 1, 2))
        Gn = np.einsum("eij,ej->ei", G, nrm)
        en = np.einsum("eij,ej->ei", eps, nrm)
        umT = np.einsum("eij,ej->ei", Pn, um)
        dn = np.einsum("eai,ei->ea", dphi, nrm)
        epsv_n = 0.5 * (np.einsum("eai,ek->eaki", dphi, nrm) + dn[:, :, None, None] * np.eye(2)[None, None])
        val = np.zeros(U.shape, dtype=U.dtype)
        val += (fs.a_p * p + fs.pconst)[:, None, None] * phi[:, :, None] * nrm[:, None, :]
        val -= fs.a_g * mu * phi[:, :, None] * Gn[:, None, :]
        val -= fs.a_s * 2.0 * mu * phi[:, :, None] * en[:, None, :]
        if fs.a_n != 0.0:
            enT = np.einsum("eik,ei->ek", Pn, en)
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
    """Q = int u_prev . n ds over the given facets with the facet rule (u_prev is Pk on the facet)."""
    k = degree_of(prob.cells)
    cells = prob.cells[pairs[:, 0]]
    lf = pairs[:, 1]
    X = prob.x[cells[:, :3]]
    ar = np.arange(cells.shape[0])
    va, vb = _FACET_VERTS[lf, 0], _FACET_VERTS[lf, 1]
    xa, xb, xo = X[ar, va], X[ar, vb], X[ar, lf]
    t = xb - xa
    nrm = np.stack([t[:, 1], -t[:, 0]], axis=1)          # |nrm| = length
    nrm = nrm * np.sign(np.einsum("ei,ei->e", nrm, xa - xo))[:, None]
    Uc = Un_nodal.reshape(-1, 2)[cells]
    pts, wts = prob.facet_rule
    total = 0.0
    for q in range(len(wts)):
        ref = (1.0 - pts[q]) * _VERT_REF[va] + pts[q] * _VERT_REF[vb]
        for lfv in range(3):
            sel = lf == lfv
            if not sel.any():
                continue
            ph, _, _ = tabulate(k, ref[sel][:1])
            uq = np.einsum("a,eai->ei", ph[0], Uc[sel])
            total += wts[q] * float(np.sum(np.einsum("ei,ei->e", uq, nrm[sel])))
    return total
