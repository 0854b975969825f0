"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  **PARITY UNPINNED** (see ns_oracle.py).

SURVEY.md §8(f) rank 4: the cell integrals of the curl-curl ("rotational") formulation of
src/solvers/stabilized_schur_pressurebc.py:85-160 on P1–P1 simplices (d = 2 triangles, d = 3 tetrahedra),
FFCx-style (full integrand at every quadrature point):

    F_u(v) = rho (v, (u - u_n)/dt) + mu (curl u_m, curl v) - (p, div v) + rho ((curl u_m x u_m) . v)
             - rho/2 (u_m . u_m, div v) - rho (v, f)                                   (:126-132)
           + (tau R, (u_m . grad) v) + tau_lsic rho (div u_m, div v)                   (:152-158)
    F_p(q) = (q, div u_m) + (1/rho) (tau R, grad q)                                    (:133, :153)
    R      = rho ((u - u_n)/dt + curl u_m x u_m) + grad p - rho f                      (:143-145)

with u_m = (u + u_n)/2 (:89), tau / tau_lsic as in stabilized_schur.py (:147-157, functions of u_n only).
In 2-D, curl u is the scalar omega = d_x u_y - d_y u_x and curl u x w = (-omega w_y, omega w_x) (:96-110).
The CUDA kernels of csrc/assembly_curlcurl.cu are checked against this oracle (tests/test_gpu_curlcurl.py); the Jacobian is the complex-step derivative of the residual (exact to
round-off: the residual is analytic in (U, P), tau depends on u_n only).
Layout as in simplex_oracle: U, Un (E, d+1, d), P (E, d+1) -> Fu (E, d+1, d), Fp (E, d+1).
"""
from __future__ import annotations

import numpy as np

from .simplex_oracle import _phi, _tau, simplex_geometry


def _curl(G, d):
    """curl of a field with gradient G[e, i, j] = d_i u_j: (E,) in 2-D, (E, 3) in 3-D."""
    if d == 2:
        return G[:, 0, 1] - G[:, 1, 0]
    return np.stack([G[:, 1, 2] - G[:, 2, 1], G[:, 2, 0] - G[:, 0, 2], G[:, 0, 1] - G[:, 1, 0]], axis=1)


def _curl_cross(om, w, d):
    """curl u x w."""
    if d == 2:
        return np.stack([-om * w[:, 1], om * w[:, 0]], axis=1)
    return np.cross(om, w)


def element_F(x, cells, h, U, P, Un, rule, dt, rho, mu, f, eps0):
    det, dphi = simplex_geometry(x, cells)
    d = x.shape[1]
    nv = d + 1
    pts, wts = rule
    f = np.asarray(f, dtype=np.float64)
    Um = 0.5 * (U + Un)
    G = np.einsum("eai,eaj->eij", dphi, Um)
    gradp = np.einsum("eai,ea->ei", dphi, P)
    divu = np.einsum("eii->e", G)
    om = _curl(G, d)
    # curl of the test functions v = phi_a e_k: gradient d_i v_j = dphi[a, i] delta_jk
    E = U.shape[0]
    curl_v = np.zeros((E, nv, d) + (() if d == 2 else (3,)), dtype=np.float64)
    for a in range(nv):
        for k in range(d):
            Gv = np.zeros((E, d, d))
            Gv[:, :, k] = dphi[:, a, :]
            curl_v[:, a, k] = _curl(Gv, d)
    cc = om[:, None, None] * curl_v if d == 2 else np.einsum("ec,eakc->eak", om, curl_v)     # curl u_m . curl v
    Fu = np.zeros(U.shape, dtype=np.result_type(U, P))
    Fp = np.zeros(P.shape, dtype=Fu.dtype)
    for q in range(len(wts)):
        phi = _phi(pts[q])
        w = wts[q] * det
        u = np.einsum("a,eai->ei", phi, U)
        un = np.einsum("a,eai->ei", phi, Un)
        um = 0.5 * (u + un)
        p = np.einsum("a,ea->e", phi, P)
        tau, tau_l = _tau(un.real, h, dt, rho, mu, eps0)
        dudt = (u - un) / dt
        rot = _curl_cross(om, um, d)
        R = rho * (dudt + rot) + gradp - rho * f[None, :]
        um_dphi = np.einsum("ei,eai->ea", um, dphi)
        half_u2 = 0.5 * np.einsum("ei,ei->e", um, um)
        Fu = Fu + w[:, None, None] * (
            rho * phi[None, :, None] * (dudt + rot - f[None, :])[:, None, :]
            + mu * cc
            - (p + rho * half_u2)[:, None, None] * dphi                   # -(p + rho |u_m|^2 / 2) div v
            + tau[:, None, None] * um_dphi[:, :, None] * R[:, None, :]
            + (tau_l * rho * divu)[:, None, None] * dphi)
        Fp = Fp + w[:, None] * (phi[None, :] * divu[:, None] + (tau / rho)[:, None] * np.einsum("ei,eai->ea", R, dphi))
    return Fu, Fp


def element_J(x, cells, h, U, P, Un, rule_rows_u, rule_rows_p, dt, rho, mu, f, eps0):
    """(E, (d+1)(d+1), (d+1)(d+1)) derivative of [Fu (rule_rows_u) ; Fp (rule_rows_p)] with respect to
    [U flattened ; P] by complex steps."""
    E, nv, d = U.shape
    nl = (d + 1) * nv
    J = np.zeros((E, nl, nl))
    step = 1e-30
    for j in range(nl):
        Uc, Pc = U.astype(complex), P.astype(complex)
        if j < d * nv:
            Uc[:, j // d, j % d] += 1j * step
        else:
            Pc[:, j - d * nv] += 1j * step
        Fu, _ = element_F(x, cells, h, Uc, Pc, Un, rule_rows_u, dt, rho, mu, f, eps0)
        _, Fp = element_F(x, cells, h, Uc, Pc, Un, rule_rows_p, dt, rho, mu, f, eps0)
        J[:, :d * nv, j] = Fu.reshape(E, -1).imag / step
        J[:, d * nv:, j] = Fp.imag / step
    return J


# --------------------------------------------------------------------------
# exterior-facet terms of stabilized_schur_pressurebc.setup (:189-205): weak pressure p_c (n.v) and the
# Nitsche terms for u_T = 0 written with curl x n,
#     pconst (n.v) + a_n [ -mu (curl u_m x n).v_T - mu (curl v x n).u_T + (beta_n mu / h) u_T.v_T ]
# (the reference passes p_inlet / 2 and p_outlet / 2 as p_c, :63-64).  (curl w x n)_k = sum_i n_i W_ik(w).
# --------------------------------------------------------------------------
def facet_F(x, cells, h, pairs, coef, U, P, Un, facet_rule, mu):
    """Fu (m, d+1, d) of one facet set; `coef` has pconst, a_n, beta_n.  Independent of P."""
    from .simplex_oracle import _facet_phi, facet_geometry
    d = x.shape[1]
    _, dphi_all = simplex_geometry(x, cells)
    dphi = dphi_all[pairs[:, 0]]
    hh = h[pairs[:, 0]]
    nrm, scale = facet_geometry(x, cells, pairs)
    Um = 0.5 * (U + Un)
    G = np.einsum("eai,eaj->eij", dphi, Um)
    W = G - np.swapaxes(G, 1, 2)
    Wn = np.einsum("ei,eik->ek", nrm, W)                   # curl u_m x n
    Pn = np.eye(d)[None] - nrm[:, :, None] * nrm[:, None, :]
    dn = np.einsum("eai,ei->ea", dphi, nrm)
    Fu = np.zeros(U.shape, dtype=np.result_type(U, P))
    pts, wts = facet_rule
    for q in range(len(wts)):
        phi = _facet_phi(d, pairs[:, 1], pts[q])
        w = wts[q] * scale
        um = np.einsum("ea,eai->ei", phi, Um)
        uT = np.einsum("eij,ej->ei", Pn, um)
        val = coef.pconst * phi[:, :, None] * nrm[:, None, :]
        WnT = np.einsum("ej,ejk->ek", Wn, Pn)
        val = val - coef.a_n * mu * phi[:, :, None] * WnT[:, None, :]
        # (curl v x n)_j for v = phi_a e_k: dn_a delta_jk - d_j phi_a n_k   (phi-independent: v enters through its gradient)
        g_uT = np.einsum("eai,ei->ea", dphi, uT)
        val = val - coef.a_n * mu * (dn[:, :, None] * uT[:, None, :] - g_uT[:, :, None] * nrm[:, None, :])
        val = val + coef.a_n * (coef.beta_n * mu / hh)[:, None, None] * phi[:, :, None] * uT[:, None, :]
        Fu = Fu + w[:, None, None] * val
    return Fu


# --------------------------------------------------------------------------
# Adapters: the element routines above behind the kernel interfaces of the global oracles, so that
# their assembly, Dirichlet treatment (assemble_*_block semantics) and Newton drivers serve the curl-curl
# formulation unchanged.  Selected with `prob.formulation = "curlcurl"`.
# --------------------------------------------------------------------------
def _blocks(J, d, nv):
    E = J.shape[0]
    nu = d * nv
    return (J[:, :nu, :nu].reshape(E, nu, nu), J[:, :nu, nu:].reshape(E, nu, nv), J[:, nu:, :nu].reshape(E, nv, nu),
            J[:, nu:, nu:].reshape(E, nv, nv))


class Kernels2D:
    """Interface of oracle/ns_oracle.py:_kernels (P1 triangles)."""

    @staticmethod
    def _check(prob, Uh):
        assert prob.theta == 0.5 and prob.a0 == 1.0 and Uh is None, "curl-curl form: default time scheme only"

    @staticmethod
    def element_F(prob, U, P, Un, rule, Uh=None):
        Kernels2D._check(prob, Uh)
        return element_F(prob.x, prob.cells, prob.h, U, P, Un, rule, prob.dt, prob.rho, prob.mu, prob.f, prob.eps0)

    @staticmethod
    def element_J(prob, U, P, Un, rule, Uh=None):
        Kernels2D._check(prob, Uh)
        J = element_J(prob.x, prob.cells, prob.h, U, P, Un, rule, rule, prob.dt, prob.rho, prob.mu, prob.f, prob.eps0)
        return _blocks(J, 2, 3)

    @staticmethod
    def facet_F(prob, fs, U, P, Un):
        return facet_F(prob.x, prob.cells, prob.h, fs.pairs, fs, U, P, Un, prob.facet_rule, prob.mu)


class KernelsSimplex:
    """Interface of oracle/simplex_oracle.py as oracle/ns3d_oracle.py calls it (any d)."""

    @staticmethod
    def element_F(x, cells, h, U, P, Un, rule, Uh=None, *, dt, rho, mu, f, eps0, theta=0.5, a0=1.0):
        assert theta == 0.5 and a0 == 1.0 and Uh is None, "curl-curl form: default time scheme only"
        return element_F(x, cells, h, U, P, Un, rule, dt, rho, mu, f, eps0)

    @staticmethod
    def element_J(x, cells, h, U, P, Un, rule, Uh=None, *, dt, rho, mu, f, eps0, theta=0.5, a0=1.0):
        assert theta == 0.5 and a0 == 1.0 and Uh is None, "curl-curl form: default time scheme only"
        d = x.shape[1]
        return _blocks(element_J(x, cells, h, U, P, Un, rule, rule, dt, rho, mu, f, eps0), d, d + 1)

    @staticmethod
    def facet_F(x, cells, h, pairs, fs, U, P, Un, facet_rule, rho, mu, theta=0.5):
        return facet_F(x, cells, h, pairs, fs, U, P, Un, facet_rule, mu)

    @staticmethod
    def facet_J(x, cells, h, pairs, fs, Un, facet_rule, rho, mu, theta=0.5):
        """(m, d nv, (d+1) nv): the facet terms are affine in (U, P) — exact derivative from unit vectors."""
        d = x.shape[1]
        nv = d + 1
        m = pairs.shape[0]
        Z2, Z1 = np.zeros((m, nv, d)), np.zeros((m, nv))
        F0 = facet_F(x, cells, h, pairs, fs, Z2, Z1, Un, facet_rule, mu)
        out = np.zeros((m, d * nv, (d + 1) * nv))
        for j in range(d * nv):                  # independent of P
            U = Z2.copy()
            U[:, j // d, j % d] = 1.0
            out[:, :, j] = (facet_F(x, cells, h, pairs, fs, U, Z1, Un, facet_rule, mu) - F0).reshape(m, d * nv)
        return out
