// Dimension-generic P1–P1 simplex element routine of the curl-curl / rotational formulation of
// src/solvers/stabilized_schur_pressurebc.py:85-160 (the `vascularbc` family is built on
// it).  Kernels: assembly_curlcurl.cu (hemo_set_formulation(ctx, HEMO_FORM_CURLCURL)).  Also checked on the host
// against oracle/curlcurl_oracle.py (tests/test_curlcurl_host.py, compiled with g++: test infrastructure only).
//
// With G_ij = d_i u_mj and the antisymmetric W_ij = G_ij - G_ji both curl expressions of the form are
// dimension-independent:
//     (curl u_m x w)_k = sum_i w_i W_ik ,      curl u_m . curl(phi_a e_k) = sum_i W_ik d_i phi_a
// (2-D: W_01 = omega, :96-110; 3-D: W_ij = eps_ijc omega_c, :112-122).  On an affine P1 cell W, div u_m and
// grad p are constant; u, u_n, p and tau vary with the quadrature point, so the integrand is evaluated
// at every point of the block form's rule (FFCx-style; a moment factorisation like simplex_element.cuh
// is the optimisation to do when this becomes a kernel).
//
//   F_u[a][k] = int rho phi_a (dudt + rot - f)_k + mu sum_i W_ik d_i phi_a - (p + rho |u_m|^2 / 2) d_k phi_a
//                   + tau (u_m . grad phi_a) R_k + tau_lsic rho div(u_m) d_k phi_a
//   F_p[a]    = int phi_a div(u_m) + (tau / rho) R . grad phi_a
//   R = rho (dudt + rot) + grad p - rho f,  dudt = (u - u_n) / dt,  rot = curl u_m x u_m,  u_m = (u + u_n) / 2
#pragma once
#include "simplex_element.cuh"

// Which block forms a rule integrates (each block form of form(extract_blocks(..)) has its own rule,
// stabilized_schur_pressurebc.py:205-206).
enum { CC_FU = 1, CC_FP = 2, CC_UU = 4, CC_UP = 8, CC_PU = 16, CC_PP = 32 };

// Residual and Jacobian of one cell with one rule.  `c` needs g, detJ, U, N, P, h, fbody
// (simplex_geometry + nodal values; simplex_derive is not used).  Outputs are ADDED:
//   Fu[a][k], Fp[a];  J[r * NL + s] with s in the cell-local order (b*D + l for velocity, D*NV + b for pressure).
// `mask` (CC_*) selects the block forms this rule integrates.  `a_only` < 0: every test node, rows r in the
// cell-local order too (J is NL x NL); `a_only` = a: only the rows of test node a, r = k for its velocity
// components and r = D for its pressure row (J is (D+1) x NL) — one thread per (cell, test node) on the device.
template <int D>
HEMO_HD void curlcurl_cell_ex(const SimplexCell<D>& c, const HemoForm& par, const SimplexRule<D>& r, int mask,
                              int a_only, double Fu[D + 1][D], double Fp[D + 1], double* J) {
    constexpr int NV = D + 1, NL = (D + 1) * (D + 1), PO = D * NV;
    const double rho = par.rho, mu = par.mu, idt = 1.0 / par.dt, th = 0.5;
    const bool rows_u = (mask & (CC_FU | CC_UU | CC_UP)) != 0, rows_p = (mask & (CC_FP | CC_PU | CC_PP)) != 0;
    const bool want_jac = (mask & (CC_UU | CC_UP | CC_PU | CC_PP)) != 0;
    double G[D][D], W[D][D], gp[D], divu = 0.0;
    for (int i = 0; i < D; ++i) {
        for (int j = 0; j < D; ++j) {
            double v = 0.0;
            for (int a = 0; a < NV; ++a) v += c.g[a][i] * 0.5 * (c.U[a][j] + c.N[a][j]);
            G[i][j] = v;
        }
        double v = 0.0;
        for (int a = 0; a < NV; ++a) v += c.g[a][i] * c.P[a];
        gp[i] = v;
    }
    for (int i = 0; i < D; ++i) {
        divu += G[i][i];
        for (int j = 0; j < D; ++j) W[i][j] = G[i][j] - G[j][i];
    }
    double gg[NV][NV];
    for (int a = 0; a < NV; ++a)
        for (int b = 0; b < NV; ++b) {
            double v = 0.0;
            for (int i = 0; i < D; ++i) v += c.g[a][i] * c.g[b][i];
            gg[a][b] = v;
        }
    const double h = c.h, nu = mu / rho;
    for (int q = 0; q < r.nq; ++q) {
        const double* phi = r.phi[q];
        const double w = r.w[q] * c.detJ;
        double u[D], un[D], um[D], p = 0.0, v2 = 0.0, um2 = 0.0;
        for (int k = 0; k < D; ++k) {
            double a1 = 0.0, a2 = 0.0;
            for (int a = 0; a < NV; ++a) { a1 += phi[a] * c.U[a][k]; a2 += phi[a] * c.N[a][k]; }
            u[k] = a1; un[k] = a2; um[k] = 0.5 * (a1 + a2);
            v2 += a2 * a2; um2 += um[k] * um[k];
        }
        for (int a = 0; a < NV; ++a) p += phi[a] * c.P[a];
        // tau_supg, tau_lsic (functions of u_n only; stabilized_schur_pressurebc.py:147-157)
        const double vn = sqrt(v2);
        const double two_v = 2.0 * vn;
        const double t1 = h / (two_v >= par.eps0 ? two_v : par.eps0);
        const double t2 = par.dt / 2.0, t3 = h * h / (4.0 * nu);
        const double tau = 1.0 / sqrt(1.0 / (t1 * t1) + 1.0 / (t2 * t2) + 1.0 / (t3 * t3));
        const double Re = vn * h / (2.0 * nu);
        const double tl = 0.5 * vn * h * (Re <= 3.0 ? Re / 3.0 : 1.0);
        double rot[D], R[D], acc[D], s[NV];
        for (int k = 0; k < D; ++k) {
            double v = 0.0;
            for (int i = 0; i < D; ++i) v += um[i] * W[i][k];
            rot[k] = v;
            acc[k] = (u[k] - un[k]) * idt + v - c.fbody[k];
            R[k] = rho * acc[k] + gp[k];
        }
        for (int a = 0; a < NV; ++a) {
            double v = 0.0;
            for (int i = 0; i < D; ++i) v += um[i] * c.g[a][i];
            s[a] = v;                                         // u_m . grad phi_a
        }
        const double bern = p + 0.5 * rho * um2;
        for (int a = 0; a < NV; ++a) {
            if (a_only >= 0 && a != a_only) continue;
            if (mask & CC_FU)
                for (int k = 0; k < D; ++k) {
                    double cc = 0.0;
                    for (int i = 0; i < D; ++i) cc += W[i][k] * c.g[a][i];
                    Fu[a][k] += w * (rho * phi[a] * acc[k] + mu * cc - bern * c.g[a][k] + tau * s[a] * R[k] +
                                     tl * rho * divu * c.g[a][k]);
                }
            if (mask & CC_FP) {
                double rg = 0.0;
                for (int i = 0; i < D; ++i) rg += R[i] * c.g[a][i];
                Fp[a] += w * (phi[a] * divu + tau / rho * rg);
            }
        }
        if (!want_jac) continue;
        for (int b = 0; b < NV; ++b) {
            // d rot_k / dU[b][l] = th (phi_b W_lk + s_b delta_kl - g_b[k] um_l);  dR_k/dU[b][l] = rho (phi_b/dt delta_kl + d rot)
            double dR[D][D];                                  // [k][l]
            for (int k = 0; k < D; ++k)
                for (int l = 0; l < D; ++l) {
                    const double dkl = (k == l) ? 1.0 : 0.0;
                    const double drot = th * (phi[b] * W[l][k] + s[b] * dkl - c.g[b][k] * um[l]);
                    dR[k][l] = rho * (phi[b] * idt * dkl + drot);
                }
            for (int a = 0; a < NV; ++a) {
                if (a_only >= 0 && a != a_only) continue;
                if (rows_u)
                    for (int k = 0; k < D; ++k) {
                        const int row = (a_only >= 0) ? k : a * D + k;
                        if (mask & CC_UU)
                            for (int l = 0; l < D; ++l) {
                                const double dkl = (k == l) ? 1.0 : 0.0;
                                double v = phi[a] * dR[k][l];                                         // rho phi_a d(acc_k)
                                v += mu * th * (gg[a][b] * dkl - c.g[b][k] * c.g[a][l]);              // curl-curl
                                v -= rho * th * um[l] * phi[b] * c.g[a][k];                           // -rho/2 |u_m|^2 div v
                                v += tau * (th * phi[b] * c.g[a][l] * R[k] + s[a] * dR[k][l]);        // SUPG
                                v += tl * rho * th * c.g[b][l] * c.g[a][k];                           // LSIC
                                J[row * NL + b * D + l] += w * v;
                            }
                        if (mask & CC_UP) J[row * NL + PO + b] += w * (-phi[b] * c.g[a][k] + tau * s[a] * c.g[b][k]);
                    }
                if (rows_p) {
                    const int row = (a_only >= 0) ? D : PO + a;
                    if (mask & CC_PU)
                        for (int l = 0; l < D; ++l) {
                            double v = phi[a] * th * c.g[b][l];
                            for (int i = 0; i < D; ++i) v += tau / rho * dR[i][l] * c.g[a][i];
                            J[row * NL + b * D + l] += w * v;
                        }
                    if (mask & CC_PP) J[row * NL + PO + b] += w * tau / rho * gg[a][b];
                }
            }
        }
    }
}

// rows_u / rows_p: the rule integrates F_u, J_uu, J_up resp. F_p, J_pu, J_pp (the host checks' calling convention)
template <int D>
HEMO_HD void curlcurl_cell(const SimplexCell<D>& c, const HemoForm& par, const SimplexRule<D>& r, bool rows_u,
                           bool rows_p, bool want_jac, double Fu[D + 1][D], double Fp[D + 1],
                           double* J /* [(D+1)^2][(D+1)^2] or nullptr */) {
    int mask = 0;
    if (rows_u) mask |= CC_FU | (want_jac ? (CC_UU | CC_UP) : 0);
    if (rows_p) mask |= CC_FP | (want_jac ? (CC_PU | CC_PP) : 0);
    curlcurl_cell_ex<D>(c, par, r, mask, -1, Fu, Fp, J);
}

// Exterior-facet terms of stabilized_schur_pressurebc.setup (:189-201) on local facet lf: weak pressure
// pconst (n.v) and the Nitsche terms for u_T = 0 written with curl x n,
//     a_n [ -mu (curl u_m x n).v_T - mu (curl v x n).u_T + (beta_n mu / h) u_T.v_T ],   (curl w x n)_k = sum_i n_i W_ik(w).
// Uses co.pconst, co.a_n, co.beta_n.  res(a, k, value) adds to F_u[a][k]; jac(a, b, k, l, value) to
// dF_u[a][k] / dU[b][l] (no pressure dependence).  `c` needs g, detJ, U, N, h.
template <int D, bool WANT_RES, bool WANT_JAC, typename Res, typename Jac>
HEMO_HD void curlcurl_facet(const SimplexCell<D>& c, const HemoForm& par, const hemo_facet_coef& co,
                            const SimplexFacetRule<D>& fr, int lf, Res res, Jac jac) {
    constexpr int NV = D + 1;
    const double mu = par.mu, th = 0.5;
    double nr[D], scale;
    simplex_facet_normal<D>(c, lf, nr, scale);
    double Pn[D][D];
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) Pn[i][j] = ((i == j) ? 1.0 : 0.0) - nr[i] * nr[j];
    double Ph[NV], Ph2[NV][NV];
    for (int a = 0; a < NV; ++a) {
        Ph[a] = 0.0;
        for (int b = 0; b < NV; ++b) Ph2[a][b] = 0.0;
    }
    for (int q = 0; q < fr.nq; ++q) {
        double phi[NV];
        for (int j = 0, v = 0; v < NV; ++v) phi[v] = (v == lf) ? 0.0 : fr.lam[q][j++];
        const double w = fr.w[q] * scale;
        for (int a = 0; a < NV; ++a) {
            Ph[a] += w * phi[a];
            for (int b = 0; b < NV; ++b) Ph2[a][b] += w * phi[a] * phi[b];
        }
    }
    double M[NV][D], dn[NV], Png[NV][D];
    for (int a = 0; a < NV; ++a) {
        dn[a] = 0.0;
        for (int i = 0; i < D; ++i) { M[a][i] = 0.5 * (c.U[a][i] + c.N[a][i]); dn[a] += c.g[a][i] * nr[i]; }
        for (int k = 0; k < D; ++k) {
            double v = 0.0;
            for (int i = 0; i < D; ++i) v += Pn[k][i] * c.g[a][i];
            Png[a][k] = v;
        }
    }
    const double pen = co.a_n * co.beta_n * mu / c.h;
    if (WANT_RES) {
        double Wn[D], WnT[D], Mi[D], MiT[D];
        for (int k = 0; k < D; ++k) {
            double v = 0.0;                                   // (curl u_m x n)_k = sum_i n_i (G_ik - G_ki)
            for (int i = 0; i < D; ++i)
                for (int a = 0; a < NV; ++a) v += nr[i] * (c.g[a][i] * M[a][k] - c.g[a][k] * M[a][i]);
            Wn[k] = v;
            double m = 0.0;
            for (int a = 0; a < NV; ++a) m += Ph[a] * M[a][k];
            Mi[k] = m;
        }
        for (int k = 0; k < D; ++k) {
            double v = 0.0, m = 0.0;
            for (int j = 0; j < D; ++j) { v += Wn[j] * Pn[j][k]; m += Pn[k][j] * Mi[j]; }
            WnT[k] = v; MiT[k] = m;
        }
        for (int a = 0; a < NV; ++a) {
            double gM = 0.0, Ma[D];
            for (int k = 0; k < D; ++k) {
                gM += c.g[a][k] * MiT[k];
                double m = 0.0;
                for (int b = 0; b < NV; ++b) m += Ph2[a][b] * M[b][k];
                Ma[k] = m;
            }
            for (int k = 0; k < D; ++k) {
                double MaT = 0.0;
                for (int i = 0; i < D; ++i) MaT += Pn[k][i] * Ma[i];
                double v = co.pconst * Ph[a] * nr[k];
                v -= co.a_n * mu * Ph[a] * WnT[k];
                v -= co.a_n * mu * (dn[a] * MiT[k] - gM * nr[k]);
                v += pen * MaT;
                res(a, k, v);
            }
        }
    }
    if (WANT_JAC) {
        for (int a = 0; a < NV; ++a)
            for (int b = 0; b < NV; ++b)
                for (int k = 0; k < D; ++k)
                    for (int l = 0; l < D; ++l) {
                        double v = -co.a_n * mu * th * Ph[a] * (dn[b] * Pn[l][k] - Png[b][k] * nr[l]);
                        v -= co.a_n * mu * th * Ph[b] * (dn[a] * Pn[k][l] - Png[a][l] * nr[k]);
                        v += th * pen * Pn[k][l] * Ph2[a][b];
                        jac(a, b, k, l, v);
                    }
    }
}
