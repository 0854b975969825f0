"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  **PARITY UNPINNED** (see ns_oracle.py).

Dimension-generic P1–P1 simplex restatement (triangles d = 2, tetrahedra d = 3) of the cell
integrals of src/solvers/stabilized_schur.py:60-123 — groundwork for the tetrahedral kernels
named by the north star (the reference runs them through `mesh.topology.cell_name()`, e.g.
src/scenarios/taylor_green.py:34 on 32^3 x 6 tetrahedra).  For d = 2 it reproduces
`ns_oracle.element_F/J` exactly (tests/test_simplex_oracle.py), which pins the generic code to the
one the CUDA path is checked against; for d = 3 it is checked against the literal sympy
transcription of the form (`form_mirror.SimplexForms`) and by complex-step differentiation.

Layout: U, Un (E, d+1, d) nodal velocities per cell, P (E, d+1); outputs Fu (E, d+1, d), Fp (E, d+1),
Juu (E, d+1, d, d+1, d) [a,k ; b,l], Jup (E, d+1, d, d+1), Jpu (E, d+1, d+1, d), Jpp (E, d+1, d+1).
"""
from __future__ import annotations

import numpy as np


def simplex_geometry(x, cells):
    """|det J| and physical gradients of the P1 basis {1 - sum xi, xi_1, ..., xi_d}: dphi (E, d+1, d)."""
    X = x[cells]                                           # (E, d+1, d)
    d = X.shape[2]
    J = np.stack([X[:, j + 1] - X[:, 0] for j in range(d)], axis=2)      # J[:, i, j] = dx_i / dxi_j
    inv = np.linalg.inv(J)
    ghat = np.vstack([-np.ones((1, d)), np.eye(d)])       # reference gradients
    dphi = np.einsum("aj,eji->eai", ghat, inv)
    return np.abs(np.linalg.det(J)), dphi


def cell_diameter(x, cells):
    """mesh.h: longest edge."""
    X = x[cells]
    nv = cells.shape[1]
    h = np.zeros(cells.shape[0])
    for i in range(nv):
        for j in range(i + 1, nv):
            h = np.maximum(h, np.linalg.norm(X[:, i] - X[:, j], axis=1))
    return h


def _tau(un, h, dt, rho, mu, eps0):
    nu = mu / rho
    vnorm = np.sqrt(np.einsum("ei,ei->e", un, un))
    two_v = 2.0 * vnorm
    t1 = h / np.where(two_v >= eps0, two_v, eps0)
    tau = (1.0 / t1 ** 2 + 1.0 / (dt / 2.0) ** 2 + 1.0 / ((h * h) / (4.0 * nu)) ** 2) ** (-0.5)
    Re = (vnorm * h) / (2.0 * nu)
    z = np.where(Re <= 3.0, Re / 3.0, 1.0)
    return tau, (vnorm * h * z) / 2.0


def _phi(pt):
    return np.concatenate([[1.0 - np.sum(pt)], pt])


def element_F(x, cells, h, U, P, Un, rule, dt, rho, mu, f, eps0, theta=0.5, a0=1.0, Uh=None):
    det, dphi = simplex_geometry(x, cells)
    d = x.shape[1]
    pts, wts = rule
    f = np.asarray(f, dtype=np.float64)
    Uh = Un if Uh is None else Uh
    Um = theta * U + (1.0 - theta) * Un
    G = np.einsum("eai,eaj->eij", dphi, Um)                # d_i u_mj
    gradp = np.einsum("eai,ea->ei", dphi, P)
    divu = np.einsum("eii->e", G)
    eps = 0.5 * (G + np.swapaxes(G, 1, 2))
    Fu = np.zeros(U.shape, dtype=U.dtype)
    Fp = np.zeros(P.shape, dtype=U.dtype)
    I = np.eye(d)
    for q in range(len(wts)):
        phi = _phi(pts[q])
        w = wts[q] * det
        u = np.einsum("a,eai->ei", phi, U)
        un = np.einsum("a,eai->ei", phi, Un)
        um = theta * u + (1.0 - theta) * un
        p = np.einsum("a,ea->e", phi, P)
        tau, tau_l = _tau(un.real, h, dt, rho, mu, eps0)
        conv = np.einsum("ei,eij->ej", um, G)
        dudt = (a0 * u - np.einsum("a,eai->ei", phi, Uh)) / dt
        sigma = 2.0 * mu * eps - p[:, None, None] * I[None]
        R = rho * (dudt + conv) + gradp - rho * f[None, :]       # div sigma = -grad p on P1 simplices
        um_dphi = np.einsum("ei,eai->ea", um, dphi)
        Fu += w[:, None, None] * (
            rho * phi[None, :, None] * (dudt + conv - f[None, :])[:, None, :]
            + np.einsum("eai,eik->eak", dphi, sigma)
            + tau[:, None, None] * um_dphi[:, :, None] * R[:, None, :]
            + (tau_l * rho * divu)[:, None, None] * dphi)
        Fp += w[:, None] * (phi[None, :] * divu[:, None]
                            + (tau / rho)[:, None] * np.einsum("ei,eai->ea", R, dphi))
    return Fu, Fp


def element_J(x, cells, h, U, P, Un, rule, dt, rho, mu, f, eps0, theta=0.5, a0=1.0, Uh=None):
    det, dphi = simplex_geometry(x, cells)
    d = x.shape[1]
    nv = d + 1
    pts, wts = rule
    f = np.asarray(f, dtype=np.float64)
    E = cells.shape[0]
    Uh = Un if Uh is None else Uh
    Um = theta * U + (1.0 - theta) * Un
    G = np.einsum("eai,eaj->eij", dphi, Um)
    gradp = np.einsum("eai,ea->ei", dphi, P)
    I = np.eye(d)
    Juu = np.zeros((E, nv, d, nv, d))
    Jup = np.zeros((E, nv, d, nv))
    Jpu = np.zeros((E, nv, nv, d))
    Jpp = np.zeros((E, nv, nv))
    dd = np.einsum("eai,ebi->eab", dphi, dphi)
    visc = theta * mu * (dd[:, :, None, :, None] * I[None, None, :, None, :]
                         + np.einsum("eal,ebk->eakbl", dphi, dphi))
    for q in range(len(wts)):
        phi = _phi(pts[q])
        w = wts[q] * det
        u = np.einsum("a,eai->ei", phi, U)
        un = np.einsum("a,eai->ei", phi, Un)
        um = theta * u + (1.0 - theta) * un
        tau, tau_l = _tau(un, h, dt, rho, mu, eps0)
        conv = np.einsum("ei,eij->ej", um, G)
        R = rho * ((a0 * u - np.einsum("a,eai->ei", phi, Uh)) / dt + conv) + gradp - rho * f[None, :]
        um_dphi = np.einsum("ei,eai->ea", um, dphi)
        dR = rho * ((a0 * phi[None, :] / dt + theta * um_dphi)[:, None, :, None] * I[None, :, None, :]
                    + theta * phi[None, None, :, None] * np.swapaxes(G, 1, 2)[:, :, None, :])      # (E,k,b,l)
        Jq = np.einsum("a,ekbl->eakbl", phi, dR) + visc
        Jq = Jq + tau[:, None, None, None, None] * (
            np.einsum("ea,ekbl->eakbl", um_dphi, dR)
            + theta * np.einsum("ek,b,eal->eakbl", R, phi, dphi))
        Jq = Jq + (theta * tau_l * rho)[:, None, None, None, None] * np.einsum("eak,ebl->eakbl", dphi, dphi)
        Juu += w[:, None, None, None, None] * Jq
        Jup += w[:, None, None, None] * (
            -np.einsum("b,eak->eakb", phi, dphi)
            + tau[:, None, None, None] * np.einsum("ea,ebk->eakb", um_dphi, dphi))
        Jpu += w[:, None, None, None] * (
            theta * np.einsum("a,ebl->eabl", phi, dphi)
            + (tau / rho)[:, None, None, None] * np.einsum("ekbl,eak->eabl", dR, dphi))
        Jpp += (w * tau / rho)[:, None, None] * dd
    return Juu, Jup, Jpu, Jpp


def tet_gauss_jacobi(degree: int):
    """Collapsed Gauss–Jacobi rule on the reference tetrahedron (what Basix uses above degree 30 and a
    valid stand-in below: the Xiao–Gimbutas tables are not available here); weights sum to 1/6."""
    from scipy.special import roots_jacobi
    m = (degree + 2) // 2

    def gj(alpha):
        xx, ww = roots_jacobi(m, alpha, 0.0)
        return 0.5 * (xx + 1.0), ww / 2.0 ** (alpha + 1.0)
    p2, w2 = gj(2.0)
    p1, w1 = gj(1.0)
    p0, w0 = gj(0.0)
    pts, wts = [], []
    for i in range(m):
        for j in range(m):
            for k in range(m):
                x0 = p2[i]
                x1 = p1[j] * (1.0 - p2[i])
                x2 = p0[k] * (1.0 - p2[i]) * (1.0 - p1[j])
                pts.append((x0, x1, x2))
                wts.append(w2[i] * w1[j] * w0[k])
    return np.array(pts), np.array(wts)


# --------------------------------------------------------------------------
# exterior-facet integrals (any d): the ds terms of src/solvers/stabilized_schur.py:79 and
# src/solvers/stabilized_schur_pressure_backflow.py:189-217, coefficients as in ns_oracle.FacetSet.
# Local facet i is opposite local vertex i; its vertices are the remaining ones in
# ascending local order, and the facet rule is given on the reference facet spanned by them
# (interval [0,1]: points (m,) or (m,1); triangle {s,t >= 0, s+t <= 1}: points (m,2), weights sum 1/2).
# --------------------------------------------------------------------------
def facet_vertices(d):
    return np.array([[v for v in range(d + 1) if v != i] for i in range(d + 1)])


def facet_geometry(x, cells, pairs):
    """Unit outward normal (m, d) and physical/reference measure ratio (m,) of the facets `pairs`
    (m, 2) = (cell, local facet)."""
    d = x.shape[1]
    X = x[cells[pairs[:, 0]]]                              # (m, d+1, d)
    m = X.shape[0]
    ar = np.arange(m)
    fv = facet_vertices(d)[pairs[:, 1]]                    # (m, d)
    P0 = X[ar, fv[:, 0]]
    if d == 2:
        t = X[ar, fv[:, 1]] - P0
        nrm = np.stack([t[:, 1], -t[:, 0]], axis=1)        # |nrm| = length = length / |[0,1]|
    else:
        nrm = np.cross(X[ar, fv[:, 1]] - P0, X[ar, fv[:, 2]] - P0)     # |nrm| = 2 area = area / (1/2)
    scale = np.linalg.norm(nrm, axis=1)
    nrm = nrm / scale[:, None]
    sgn = np.sign(np.einsum("ei,ei->e", nrm, P0 - X[ar, pairs[:, 1]]))  # away from the opposite vertex
    return nrm * sgn[:, None], scale


def _facet_phi(d, lf, pt):
    """(m, d+1) cell basis at the reference-facet point `pt` of the local facets lf (m,)."""
    pt = np.atleast_1d(np.asarray(pt, dtype=np.float64))
    lam = np.concatenate([[1.0 - pt.sum()], pt])           # barycentric on the facet's own vertices
    fv = facet_vertices(d)[lf]
    phi = np.zeros((len(lf), d + 1))
    for j in range(d):
        phi[np.arange(len(lf)), fv[:, j]] = lam[j]
    return phi


def facet_F(x, cells, h, pairs, coef, U, P, Un, facet_rule, rho, mu, theta=0.5):
    """Fu (m, d+1, d) of one facet set; U, Un (m, d+1, d), P (m, d+1) are the nodal values of the
    cells pairs[:, 0].  `coef`: object with a_p, pconst, a_g, a_s, a_n, beta_n, a_b, beta_b."""
    d = x.shape[1]
    _, dphi_all = simplex_geometry(x, cells)
    dphi = dphi_all[pairs[:, 0]]
    hh = h[pairs[:, 0]]
    nrm, scale = facet_geometry(x, cells, pairs)
    Um = theta * U + (1.0 - theta) * Un
    G = np.einsum("eai,eaj->eij", dphi, Um)                # d_i u_mj = nabla_grad(u_m)
    eps = 0.5 * (G + np.swapaxes(G, 1, 2))
    Gn = np.einsum("eij,ej->ei", G, nrm)                   # dot(nabla_grad(u_m), n)   (:79)
    en = np.einsum("eij,ej->ei", eps, nrm)
    I = np.eye(d)
    Pn = I[None] - nrm[:, :, None] * nrm[:, None, :]
    dn = np.einsum("eai,ei->ea", dphi, nrm)
    epsv_n = 0.5 * (np.einsum("eai,ek->eaki", dphi, nrm) + dn[:, :, None, None] * I[None, None])
    Fu = np.zeros(U.shape, dtype=np.result_type(U, P))
    pts, wts = facet_rule
    for q in range(len(wts)):
        phi = _facet_phi(d, pairs[:, 1], pts[q])
        w = wts[q] * scale
        um = np.einsum("ea,eai->ei", phi, Um)
        un = np.einsum("ea,eai->ei", phi, Un).real
        p = np.einsum("ea,ea->e", phi, P)
        umT = np.einsum("eij,ej->ei", Pn, um)
        val = (coef.a_p * p + coef.pconst)[:, None, None] * phi[:, :, None] * nrm[:, None, :]
        val = val - coef.a_g * mu * phi[:, :, None] * Gn[:, None, :]
        val = val - coef.a_s * 2.0 * mu * phi[:, :, None] * en[:, None, :]
        if coef.a_n != 0.0:
            enT = np.einsum("eik,ei->ek", Pn, en)
            val = val - coef.a_n * 2.0 * mu * phi[:, :, None] * enT[:, None, :]
            val = val - coef.a_n * 2.0 * mu * np.einsum("eaki,ei->eak", epsv_n, umT)
            val = val + coef.a_n * (coef.beta_n * mu / hh)[:, None, None] * phi[:, :, None] * umT[:, None, :]
        if coef.a_b != 0.0:
            unn = np.einsum("ei,ei->e", un, nrm)
            un_minus = 0.5 * (unn - np.abs(unn))
            val = val - coef.a_b * coef.beta_b * rho * un_minus[:, None, None] * phi[:, :, None] * um[:, None, :]
        Fu = Fu + w[:, None, None] * val
    return Fu


def facet_J(x, cells, h, pairs, coef, Un, facet_rule, rho, mu, theta=0.5):
    """(m, d(d+1), (d+1)(d+1)) derivative of facet_F with respect to (U flattened, then P): the facet
    terms are affine in (U, P), so unit vectors give it exactly."""
    d = x.shape[1]
    nv = d + 1
    m = pairs.shape[0]
    Z2, Z1 = np.zeros((m, nv, d)), np.zeros((m, nv))
    F0 = facet_F(x, cells, h, pairs, coef, Z2, Z1, Un, facet_rule, rho, mu, theta)
    out = np.zeros((m, d * nv, (d + 1) * nv))
    for j in range((d + 1) * nv):
        U, P = Z2.copy(), Z1.copy()
        if j < d * nv:
            U[:, j // d, j % d] = 1.0
        else:
            P[:, j - d * nv] = 1.0
        out[:, :, j] = (facet_F(x, cells, h, pairs, coef, U, P, Un, facet_rule, rho, mu, theta) - F0).reshape(m, d * nv)
    return out


def outlet_flux(x, cells, pairs, un_nodal):
    """Q = int u_prev . n ds over the facets (pressure_backflow.py:204-211, 383-385); exact for P1:
    facet measure times the mean of the facet's nodal values."""
    d = x.shape[1]
    nrm, scale = facet_geometry(x, cells, pairs)
    measure = scale * (1.0 if d == 2 else 0.5)
    Uc = un_nodal.reshape(-1, d)[cells[pairs[:, 0]]]
    fv = facet_vertices(d)[pairs[:, 1]]
    ar = np.arange(pairs.shape[0])
    mean = sum(Uc[ar, fv[:, j]] for j in range(d)) / d
    return float(np.sum(np.einsum("ei,ei->e", mean, nrm) * measure))


def exterior_facets(cells):
    """(m, 2) (cell, local facet) pairs of the facets attached to exactly one cell
    (dolfinx.mesh.exterior_facet_indices + compute_integration_domains), sorted by (cell, facet)."""
    E, nv = cells.shape
    d = nv - 1
    fv = facet_vertices(d)
    keys = np.sort(cells[:, fv], axis=2).reshape(E * nv, d)               # (E*nv, d) sorted vertex tuples
    _, inv, cnt = np.unique(keys, axis=0, return_inverse=True, return_counts=True)
    ext = np.nonzero(cnt[inv.reshape(-1)] == 1)[0]
    return np.stack([ext // nv, ext % nv], axis=1).astype(np.int32)


def triangle_facet_rule(degree=2):
    """Symmetric rules on the reference triangle for the tetrahedron's facets (weights sum 1/2):
    degree 2 -> 3 interior points, degree 3/4 -> 6 points (Dunavant)."""
    if degree <= 2:
        pts = np.array([[1 / 6, 1 / 6], [2 / 3, 1 / 6], [1 / 6, 2 / 3]])
        return pts, np.full(3, 1 / 6)
    a, b = 0.445948490915965, 0.091576213509771
    wa, wb = 0.223381589678011, 0.109951743655322
    pts = np.array([[a, a], [1 - 2 * a, a], [a, 1 - 2 * a], [b, b], [1 - 2 * b, b], [b, 1 - 2 * b]])
    return pts, 0.5 * np.array([wa, wa, wa, wb, wb, wb])
