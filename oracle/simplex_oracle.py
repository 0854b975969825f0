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
