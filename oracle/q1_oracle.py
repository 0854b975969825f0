"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  **PARITY UNPINNED** (see ns_oracle.py).

Q1–Q1 quadrilateral restatement of the same forms as `ns_oracle.py`
(reference src/solvers/stabilized_schur.py:60-123 with
`mesh.topology.cell_name() == "quadrilateral"`, which is what
src/scenarios/stenosis_pressure_structured.py:379-386 produces through
`setTransfiniteSurface` + `setRecombine`).  Differences from P1 triangles:

  * tensor-ordered vertices (0,0),(1,0),(0,1),(1,1); basis
    {(1-x)(1-y), x(1-y), (1-x)y, xy}; facets (0,1),(0,2),(1,3),(2,3)
    (3P Basix numbering);
  * the geometry is bilinear, so J, det J and grad(phi) vary inside the cell
    and are evaluated at every quadrature point;
  * `div(sigma(u_mid, p))` in the strong residual R (:95-97) no longer
    vanishes: UFL differentiates K = J^-1 for non-affine cells, i.e. the exact
    physical Hessian  H(phi_a) = theta_a * kappa  with
        kappa_ij = K_0i K_1j + K_1i K_0j,
        theta_a  = d2phi_a/dxi deta - grad(phi_a) . d2x/dxi deta
    (only the mixed reference derivative of a bilinear map is non-zero);
  * the default rule is the tensor Gauss–Jacobi (= Gauss–Legendre) rule Basix
    selects for non-simplex cells; the estimated degrees (SURVEY.md §7.1,
    hand-derived) are F_u, J_uu 22; F_p, J_up, J_pu 20; J_pp 18.

Element integrals are evaluated FFCx-style (full integrand at every point),
vectorised over cells.  Every function takes the same `Problem` as ns_oracle.
"""
from __future__ import annotations

import numpy as np

_S = np.array([1.0, -1.0, -1.0, 1.0])          # d2 phi_a / dxi deta
FACET_VERTS_Q = np.array([[0, 1], [0, 2], [1, 3], [2, 3]])


def basis(xi, eta):
    """phi (4,), dphi/dxi (4, 2) of the Q1 reference basis at one point."""
    phi = np.array([(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta])
    dphi = np.array([[-(1 - eta), -(1 - xi)], [(1 - eta), -xi], [-eta, (1 - xi)], [eta, xi]])
    return phi, dphi


def point_geometry(X, xi, eta):
    """Geometry at one reference point for all cells.  X: (E,4,2).
    Returns phi (4,), g (E,4,2) physical gradients, detJ (E,) (signed),
    theta (E,4), kappa (E,2,2)."""
    phi, dref = basis(xi, eta)
    J = np.einsum("eai,aj->eij", X, dref)               # J_ij = dx_i/dxi_j
    det = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
    K = np.empty_like(J)
    K[:, 0, 0] = J[:, 1, 1] / det
    K[:, 0, 1] = -J[:, 0, 1] / det
    K[:, 1, 0] = -J[:, 1, 0] / det
    K[:, 1, 1] = J[:, 0, 0] / det
    g = np.einsum("aj,eji->eai", dref, K)               # d_i phi_a = dphi_a/dxi_j K_ji
    c = np.einsum("a,eai->ei", _S, X)                    # d2x/dxi deta
    theta = _S[None, :] - np.einsum("eai,ei->ea", g, c)
    kappa = (np.einsum("ei,ej->eij", K[:, 0, :], K[:, 1, :])
             + np.einsum("ei,ej->eij", K[:, 1, :], K[:, 0, :]))
    return phi, g, det, theta, kappa


def cell_diameter(x, cells):
    """mesh.h for quadrilaterals: largest vertex-vertex distance."""
    X = x[cells]
    h = np.zeros(cells.shape[0])
    for i in range(4):
        for j in range(i + 1, 4):
            h = np.maximum(h, np.linalg.norm(X[:, i] - X[:, j], axis=1))
    return h


def _tau(prob, unq, h):
    from . import ns_oracle as O
    return O._tau(prob, unq, h)


def _point_state(prob, X, U, P, Un, xi, eta, Uh=None):
    rho, mu, dt = prob.rho, prob.mu, prob.dt
    th, a0 = prob.theta, prob.a0          # time scheme, see ns_oracle.Problem
    Uh = Un if Uh is None else Uh
    f = np.asarray(prob.f, dtype=np.float64)
    phi, g, det, theta, kappa = point_geometry(X, xi, eta)
    Um = th * U + (1.0 - th) * Un
    u = np.einsum("a,eai->ei", phi, U)
    un = np.einsum("a,eai->ei", phi, Un)
    um = th * u + (1.0 - th) * un
    p = np.einsum("a,ea->e", phi, P)
    G = np.einsum("eai,eaj->eij", g, Um)                  # d_i u_mj
    gradp = np.einsum("eai,ea->ei", g, P)
    divu = G[:, 0, 0] + G[:, 1, 1]
    eps = 0.5 * (G + np.swapaxes(G, 1, 2))
    conv = np.einsum("ei,eij->ej", um, G)
    dudt = (a0 * u - np.einsum("a,eai->ei", phi, Uh)) / dt
    trk = kappa[:, 0, 0] + kappa[:, 1, 1]
    wv = np.einsum("ea,eak->ek", theta, Um)               # sum_b theta_b M_b
    # div(2 mu eps(u_m)) = mu (lap u + grad div u)
    visc = mu * (wv * trk[:, None] + np.einsum("ekj,ej->ek", kappa, wv))
    R = rho * (dudt + conv - f[None, :]) + gradp - visc
    tau, tau_l = _tau(prob, un.real, prob.h)
    return dict(phi=phi, g=g, det=det, theta=theta, kappa=kappa, trk=trk, u=u, un=un, um=um, p=p, G=G,
                gradp=gradp, divu=divu, eps=eps, conv=conv, dudt=dudt, R=R, tau=tau, tau_l=tau_l)


def element_F(prob, U, P, Un, rule, Uh=None):
    """Element residual (Fu (E,4,2), Fp (E,4)) with the given rule (pts on [0,1]^2, wts sum 1)."""
    X = prob.x[prob.cells]
    pts, wts = rule
    rho, mu = prob.rho, prob.mu
    f = np.asarray(prob.f, dtype=np.float64)
    Fu = np.zeros(U.shape, dtype=U.dtype)
    Fp = np.zeros(P.shape, dtype=U.dtype)
    for q in range(len(wts)):
        s = _point_state(prob, X, U, P, Un, pts[q, 0], pts[q, 1], Uh)
        w = wts[q] * np.abs(s["det"])
        phi, g = s["phi"], s["g"]
        sigma = 2.0 * mu * s["eps"] - s["p"][:, None, None] * np.eye(2)[None]
        um_g = np.einsum("ei,eai->ea", s["um"], g)
        Fu += w[:, None, None] * (
            rho * phi[None, :, None] * (s["dudt"] + s["conv"] - f[None, :])[:, None, :]
            + np.einsum("eai,eik->eak", g, sigma)
            + s["tau"][:, None, None] * um_g[:, :, None] * s["R"][:, None, :]
            + (s["tau_l"] * rho * s["divu"])[:, None, None] * g)
        Fp += w[:, None] * (phi[None, :] * s["divu"][:, None]
                            + (s["tau"] / rho)[:, None] * np.einsum("ei,eai->ea", s["R"], g))
    return Fu, Fp


def element_J(prob, U, P, Un, rule, Uh=None):
    """Element Jacobian blocks: Juu (E,4,2,4,2) [a,k ; b,l], Jup (E,4,2,4),
    Jpu (E,4,4,2), Jpp (E,4,4)."""
    X = prob.x[prob.cells]
    pts, wts = rule
    rho, mu, dt = prob.rho, prob.mu, prob.dt
    th, a0 = prob.theta, prob.a0
    E = prob.cells.shape[0]
    I2 = np.eye(2)
    Juu = np.zeros((E, 4, 2, 4, 2))
    Jup = np.zeros((E, 4, 2, 4))
    Jpu = np.zeros((E, 4, 4, 2))
    Jpp = np.zeros((E, 4, 4))
    for q in range(len(wts)):
        s = _point_state(prob, X, U, P, Un, pts[q, 0], pts[q, 1], Uh)
        w = wts[q] * np.abs(s["det"])
        phi, g, G, tau, tau_l = s["phi"], s["g"], s["G"], s["tau"], s["tau_l"]
        um_g = np.einsum("ei,eai->ea", s["um"], g)
        dd = np.einsum("eai,ebi->eab", g, g)
        # C[k,b,l] = rho [ (a0 phi_b/dt + th um.g_b) d_kl + th phi_b G_lk ]   (th = d u_e / d u)
        C = rho * ((a0 * phi[None, :] / dt + th * um_g)[:, None, :, None] * I2[None, :, None, :]
                   + th * phi[None, None, :, None] * np.swapaxes(G, 1, 2)[:, :, None, :])
        # dR = C - th mu theta_b (tr(kappa) d_kl + kappa_kl)
        dR = C - th * mu * s["theta"][:, None, :, None] * (
            s["trk"][:, None, None, None] * I2[None, :, None, :] + s["kappa"][:, :, None, :])
        visc = th * mu * (dd[:, :, None, :, None] * I2[None, None, :, None, :]
                          + np.einsum("eal,ebk->eakbl", g, g))
        Jq = np.einsum("a,ekbl->eakbl", phi, C) + visc
        Jq = Jq + tau[:, None, None, None, None] * (
            np.einsum("ea,ekbl->eakbl", um_g, dR)
            + th * np.einsum("ek,b,eal->eakbl", s["R"], phi, g))
        Jq = Jq + (th * tau_l * rho)[:, None, None, None, None] * np.einsum("eak,ebl->eakbl", g, g)
        Juu += w[:, None, None, None, None] * Jq
        Jup += w[:, None, None, None] * (
            -np.einsum("b,eak->eakb", phi, g)
            + tau[:, None, None, None] * np.einsum("ea,ebk->eakb", um_g, g))
        Jpu += w[:, None, None, None] * (
            th * np.einsum("a,ebl->eabl", phi, g)
            + (tau / rho)[:, None, None, None] * np.einsum("ekbl,eak->eabl", dR, g))
        Jpp += (w * tau / rho)[:, None, None] * dd
    return Juu, Jup, Jpu, Jpp


def facet_normals(X, lf):
    """Unit outward normal and length of local facet lf of each cell (straight edges)."""
    ar = np.arange(X.shape[0])
    va, vb = FACET_VERTS_Q[lf, 0], FACET_VERTS_Q[lf, 1]
    xa, xb = X[ar, va], X[ar, vb]
    t = xb - xa
    length = np.linalg.norm(t, axis=1)
    nrm = np.stack([t[:, 1], -t[:, 0]], axis=1) / length[:, None]
    xc = X.mean(axis=1)
    sgn = np.sign(np.einsum("ei,ei->e", nrm, 0.5 * (xa + xb) - xc))
    return nrm * sgn[:, None], length, va, vb


def _facet_ref_point(lf, s):
    """Reference coordinates of the point at parameter s on local facet lf."""
    xi = np.where(lf == 0, s, np.where(lf == 1, 0.0, np.where(lf == 2, 1.0, s)))
    eta = np.where(lf == 0, 0.0, np.where(lf == 1, s, np.where(lf == 2, s, 1.0)))
    return xi, eta


def facet_F(prob, fs, U, P, Un):
    """Element residual contributions (Fu (m,4,2)) of one facet set; same terms
    as ns_oracle.facet_F with gradients evaluated at the facet points."""
    ce = fs.pairs[:, 0]
    lf = fs.pairs[:, 1]
    X = prob.x[prob.cells[ce]]
    h = prob.h[ce]
    m = X.shape[0]
    nrm, length, va, vb = facet_normals(X, lf)
    mu, rho = prob.mu, prob.rho
    Um = prob.theta * U + (1.0 - prob.theta) * Un
    Fu = np.zeros(U.shape, dtype=U.dtype)
    pts, wts = prob.facet_rule
    Pn = np.eye(2)[None] - nrm[:, :, None] * nrm[:, None, :]
    for q in range(len(wts)):
        xi, eta = _facet_ref_point(lf, pts[q])
        # per-cell reference points: evaluate the basis cell by cell (vectorised formulas)
        phi = np.stack([(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta], axis=1)       # (m,4)
        dref = np.stack([np.stack([-(1 - eta), -(1 - xi)], 1), np.stack([(1 - eta), -xi], 1),
                         np.stack([-eta, (1 - xi)], 1), np.stack([eta, xi], 1)], axis=1)                # (m,4,2)
        J = np.einsum("eai,eaj->eij", X, dref)
        det = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
        K = np.empty_like(J)
        K[:, 0, 0] = J[:, 1, 1] / det
        K[:, 0, 1] = -J[:, 0, 1] / det
        K[:, 1, 0] = -J[:, 1, 0] / det
        K[:, 1, 1] = J[:, 0, 0] / det
        g = np.einsum("eaj,eji->eai", dref, K)
        w = wts[q] * length
        G = np.einsum("eai,eaj->eij", g, Um)
        eps = 0.5 * (G + np.swapaxes(G, 1, 2))
        Gn = np.einsum("eij,ej->ei", G, nrm)
        en = np.einsum("eij,ej->ei", eps, nrm)
        dn = np.einsum("eai,ei->ea", g, nrm)
        epsv_n = 0.5 * (np.einsum("eai,ek->eaki", g, nrm) + dn[:, :, None, None] * np.eye(2)[None, None])
        um = np.einsum("ea,eai->ei", phi, Um)
        un = np.einsum("ea,eai->ei", phi, Un).real
        p = np.einsum("ea,ea->e", phi, P)
        umT = np.einsum("eij,ej->ei", Pn, um)
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
    """Q = int u_prev . n ds (pressure_backflow.py:204-211, 383-385); the trace of
    Q1 on a straight edge is linear, so the mid-point value times the length is exact."""
    X = prob.x[prob.cells[pairs[:, 0]]]
    nrm, length, va, vb = facet_normals(X, pairs[:, 1])
    Uc = Un_nodal.reshape(-1, 2)[prob.cells[pairs[:, 0]]]
    ar = np.arange(X.shape[0])
    umid = 0.5 * (Uc[ar, va] + Uc[ar, vb])
    return float(np.sum(np.einsum("ei,ei->e", umid, nrm) * length))


def tensor_gauss(m: int):
    """m x m Gauss–Legendre rule on [0,1]^2 (Basix Gauss–Jacobi scheme on a
    quadrilateral, m = (degree + 2) // 2); first coordinate slowest."""
    x, w = np.polynomial.legendre.leggauss(m)
    x = 0.5 * (x + 1.0)
    w = 0.5 * w
    pts = np.stack([np.repeat(x, m), np.tile(x, m)], axis=1)
    wts = np.repeat(w, m) * np.tile(w, m)
    return pts, wts
