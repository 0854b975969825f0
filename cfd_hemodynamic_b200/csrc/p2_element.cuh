// P2-P2 triangle element routines (host + device): the cell and exterior-facet integrals of the stabilized forms for
// Lagrange degree 2 in both spaces — what the reference builds with `p_grade = 2`
// (src/solvers/stabilized_schur_pressure_backflow.py:71,102-161,170-217; src/solvers/stabilized_schur_backflow.py:63,85-176).
//
// FFCx-style arithmetic: the full integrand at every point of each block form's own rule (rules are a run-time input).
// Affine geometry: the physical gradients follow from the constant J^-1, the physical Hessians of the six basis
// functions are constant per cell, and with them the viscous part of the strong residual
//   -div sigma(u_m, p) = -mu (Laplace u_m + grad div u_m) + grad p          ("Pk >= 2")
// and its derivative with respect to the nodal values.
//
// Local node order (3P, Basix): vertices 0, 1, 2, then the edge nodes of the edges opposite vertices 0, 1, 2.
// Element buffers use the generic layout with nv = 6: Ae[(a*6+b)*9 + ri*3+ci][E], Fe[a*3+comp][E].
// Work decomposition of the Jacobian: one work item per (cell, test node a): the three rows (u_x, u_y, p) of node a
// against all 18 columns, integrated block form by block form (at most 24 accumulators live at a time).
// Checked on the host against oracle/pk_oracle.py (tests/host_p2, tests/test_p2_host.py) and on the GPU
// (tests/test_gpu_p2.py).
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/hemo.h"
#include "hemo_rules.h"

#ifndef HEMO_HD
#ifdef __CUDACC__
#define HEMO_HD __host__ __device__ __forceinline__
#else
#define HEMO_HD inline
#endif
#endif

#define HEMO_MAXQ_P2 128        // points per cell rule (collapsed Gauss-Jacobi degree 20: 121; Basix Xiao-Gimbutas: 79)

// Cell rule of one block form on the reference triangle: (xi, eta, weight), weights sum to 1/2
struct HemoP2Rule {
    int nq;
    int alias;                  // lowest block id of the same group with an identical rule
    double pt[HEMO_MAXQ_P2][3];
};

static inline void hemo_p2_rule_aliases(HemoP2Rule* rules, const bool* have, int nrules) {
    for (int b = 0; b < nrules; ++b) {
        if (!have[b]) continue;
        HemoP2Rule& rb = rules[b];
        rb.alias = b;
        const int first = (b <= 1) ? 0 : 2;
        for (int a = first; a < b; ++a) {
            if (!have[a] || rules[a].nq != rb.nq) continue;
            bool same = true;
            for (int q = 0; q < rb.nq && same; ++q)
                same = rules[a].pt[q][0] == rb.pt[q][0] && rules[a].pt[q][1] == rb.pt[q][1] && rules[a].pt[q][2] == rb.pt[q][2];
            if (same) { rb.alias = a; break; }
        }
    }
}

struct P2Cell {
    double X[3][2];
    double U[6][2], N[6][2], H[6][2], P[6];
    double h;
    // derived (p2_prepare)
    double K[2][2];             // K[j][i] = d xi_j / d x_i
    double adet;
    double hess[6][3];          // physical Hessian of phi_b: (xx, xy, yy)
    double lapb[6];             // Laplace phi_b
    double viscR[2];            // -mu (Laplace u_m + grad div u_m): constant on the cell
    double inv_h2, c23, re_fac, half_h;
};

struct P2Point {
    double phi[6], g[6][2];
    double um[2], G[2][2], gp[2], divu, p, acc[2], R[2], tau, taul, umg[6];
};

// barycentric gradient table d lambda_a / d (xi, eta)
HEMO_HD void p2_dl(int a, double& d0, double& d1) {
    d0 = (a == 0) ? -1.0 : (a == 1 ? 1.0 : 0.0);
    d1 = (a == 0) ? -1.0 : (a == 2 ? 1.0 : 0.0);
}

HEMO_HD void p2_edge_verts(int e, int& i, int& j) {      // edge e is opposite vertex e
    i = (e == 0) ? 1 : 0;
    j = (e == 2) ? 1 : 2;
}

// reference basis and gradients at (xi, eta)
HEMO_HD void p2_tabulate(double xi, double eta, double phi[6], double dr[6][2]) {
    const double l[3] = {1.0 - xi - eta, xi, eta};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double d0, d1;
        p2_dl(a, d0, d1);
        phi[a] = l[a] * (2.0 * l[a] - 1.0);
        const double f = 4.0 * l[a] - 1.0;
        dr[a][0] = f * d0; dr[a][1] = f * d1;
    }
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        int i, j;
        p2_edge_verts(e, i, j);
        double di0, di1, dj0, dj1;
        p2_dl(i, di0, di1);
        p2_dl(j, dj0, dj1);
        phi[3 + e] = 4.0 * l[i] * l[j];
        dr[3 + e][0] = 4.0 * (l[i] * dj0 + l[j] * di0);
        dr[3 + e][1] = 4.0 * (l[i] * dj1 + l[j] * di1);
    }
}

// geometry, Hessians, stabilisation constants and the constant viscous strong-residual part; call after loading
HEMO_HD void p2_prepare(P2Cell& c, const HemoForm& par) {
    const double J00 = c.X[1][0] - c.X[0][0], J01 = c.X[2][0] - c.X[0][0];
    const double J10 = c.X[1][1] - c.X[0][1], J11 = c.X[2][1] - c.X[0][1];
    const double det = J00 * J11 - J01 * J10;
    const double id = 1.0 / det;
    c.K[0][0] = J11 * id; c.K[0][1] = -J01 * id;
    c.K[1][0] = -J10 * id; c.K[1][1] = J00 * id;
    c.adet = fabs(det);
    // physical gradient of lambda_a: gl[a][i] = sum_j dl[a][j] K[j][i]
    double gl[3][2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double d0, d1;
        p2_dl(a, d0, d1);
        gl[a][0] = d0 * c.K[0][0] + d1 * c.K[1][0];
        gl[a][1] = d0 * c.K[0][1] + d1 * c.K[1][1];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {           // phi_a = l_a (2 l_a - 1): Hessian 4 grad l_a (x) grad l_a
        c.hess[a][0] = 4.0 * gl[a][0] * gl[a][0];
        c.hess[a][1] = 4.0 * gl[a][0] * gl[a][1];
        c.hess[a][2] = 4.0 * gl[a][1] * gl[a][1];
    }
#pragma unroll
    for (int e = 0; e < 3; ++e) {           // phi = 4 l_i l_j: Hessian 4 (grad l_i (x) grad l_j + grad l_j (x) grad l_i)
        int i, j;
        p2_edge_verts(e, i, j);
        c.hess[3 + e][0] = 8.0 * gl[i][0] * gl[j][0];
        c.hess[3 + e][1] = 4.0 * (gl[i][0] * gl[j][1] + gl[j][0] * gl[i][1]);
        c.hess[3 + e][2] = 8.0 * gl[i][1] * gl[j][1];
    }
    double lap[2] = {0.0, 0.0}, gd[2] = {0.0, 0.0};
    const double th = par.theta;
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        c.lapb[b] = c.hess[b][0] + c.hess[b][2];
        const double m0 = th * c.U[b][0] + (1.0 - th) * c.N[b][0], m1 = th * c.U[b][1] + (1.0 - th) * c.N[b][1];
        lap[0] += c.lapb[b] * m0; lap[1] += c.lapb[b] * m1;
        gd[0] += c.hess[b][0] * m0 + c.hess[b][1] * m1;        // d_x (d_x u_x + d_y u_y)
        gd[1] += c.hess[b][1] * m0 + c.hess[b][2] * m1;
    }
    c.viscR[0] = -par.mu * (lap[0] + gd[0]);
    c.viscR[1] = -par.mu * (lap[1] + gd[1]);
    const double h = c.h;
    c.inv_h2 = 1.0 / (h * h);
    const double t2inv = 2.0 * par.inv_dt, t3inv = 4.0 * par.nu * c.inv_h2;
    c.c23 = t2inv * t2inv + t3inv * t3inv;
    c.re_fac = h / (2.0 * par.nu);
    c.half_h = 0.5 * h;
}

HEMO_HD void p2_tau(const HemoForm& par, const P2Cell& c, double unx, double uny, double& tau, double& taul) {
    const double v2 = unx * unx + uny * uny;
    const double t1 = fmax(4.0 * v2, par.eps0 * par.eps0) * c.inv_h2;
#ifdef __CUDA_ARCH__
    tau = rsqrt(t1 + c.c23);
#else
    tau = 1.0 / sqrt(t1 + c.c23);
#endif
    const double v = sqrt(v2);
    const double Re = v * c.re_fac;
    const double z = (Re <= 3.0) ? Re * (1.0 / 3.0) : 1.0;
    taul = c.half_h * v * z;
}

// everything the integrands need at one point
HEMO_HD void p2_point(const P2Cell& c, const HemoForm& par, double xi, double eta, P2Point& s) {
    double dr[6][2];
    p2_tabulate(xi, eta, s.phi, dr);
    double u[2] = {0, 0}, un[2] = {0, 0}, uh[2] = {0, 0};
    const double th = par.theta;
    s.p = 0.0;
    s.G[0][0] = s.G[0][1] = s.G[1][0] = s.G[1][1] = 0.0;
    s.gp[0] = s.gp[1] = 0.0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        s.g[a][0] = dr[a][0] * c.K[0][0] + dr[a][1] * c.K[1][0];
        s.g[a][1] = dr[a][0] * c.K[0][1] + dr[a][1] * c.K[1][1];
        const double m0 = th * c.U[a][0] + (1.0 - th) * c.N[a][0], m1 = th * c.U[a][1] + (1.0 - th) * c.N[a][1];
        u[0] += s.phi[a] * c.U[a][0]; u[1] += s.phi[a] * c.U[a][1];
        un[0] += s.phi[a] * c.N[a][0]; un[1] += s.phi[a] * c.N[a][1];
        uh[0] += s.phi[a] * c.H[a][0]; uh[1] += s.phi[a] * c.H[a][1];
        s.p += s.phi[a] * c.P[a];
        s.G[0][0] += s.g[a][0] * m0; s.G[0][1] += s.g[a][0] * m1;
        s.G[1][0] += s.g[a][1] * m0; s.G[1][1] += s.g[a][1] * m1;
        s.gp[0] += s.g[a][0] * c.P[a]; s.gp[1] += s.g[a][1] * c.P[a];
    }
    s.um[0] = th * u[0] + (1.0 - th) * un[0];
    s.um[1] = th * u[1] + (1.0 - th) * un[1];
    s.divu = s.G[0][0] + s.G[1][1];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const double conv = s.um[0] * s.G[0][k] + s.um[1] * s.G[1][k];
        s.acc[k] = (par.a0 * u[k] - uh[k]) * par.inv_dt + conv - par.f[k];
        s.R[k] = par.rho * s.acc[k] + c.viscR[k] + s.gp[k];
    }
    p2_tau(par, c, un[0], un[1], s.tau, s.taul);
#pragma unroll
    for (int a = 0; a < 6; ++a) s.umg[a] = s.um[0] * s.g[a][0] + s.um[1] * s.g[a][1];
}

// ---- residual -----------------------------------------------------------------------------------------------------
HEMO_HD void p2_cell_residual(const P2Cell& c, const HemoForm& par, const HemoP2Rule* rules, double Fu[6][2], double Fp[6]) {
#pragma unroll
    for (int a = 0; a < 6; ++a) { Fu[a][0] = Fu[a][1] = 0.0; Fp[a] = 0.0; }
    const bool shared = rules[HEMO_Q_FP].alias == HEMO_Q_FU;
    for (int r = HEMO_Q_FU; r <= HEMO_Q_FP; ++r) {
        if (r == HEMO_Q_FP && shared) break;
        const HemoP2Rule& ru = rules[r];
        const bool do_u = (r == HEMO_Q_FU), do_p = (r == HEMO_Q_FP) || shared;
        for (int q = 0; q < ru.nq; ++q) {
            P2Point s;
            p2_point(c, par, ru.pt[q][0], ru.pt[q][1], s);
            const double w = ru.pt[q][2] * c.adet;
            if (do_u) {
                const double e01 = 0.5 * (s.G[0][1] + s.G[1][0]);
                const double sig[2][2] = {{2.0 * par.mu * s.G[0][0] - s.p, 2.0 * par.mu * e01},
                                          {2.0 * par.mu * e01, 2.0 * par.mu * s.G[1][1] - s.p}};
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        Fu[a][k] += w * (par.rho * s.phi[a] * s.acc[k] + s.g[a][0] * sig[0][k] + s.g[a][1] * sig[1][k] +
                                         s.tau * s.umg[a] * s.R[k] + s.taul * par.rho * s.divu * s.g[a][k]);
            }
            if (do_p) {
#pragma unroll
                for (int a = 0; a < 6; ++a)
                    Fp[a] += w * (s.phi[a] * s.divu + s.tau * par.inv_rho * (s.R[0] * s.g[a][0] + s.R[1] * s.g[a][1]));
            }
        }
    }
}

// ---- Jacobian rows of one test node -------------------------------------------------------------------------------
// dR_k / dU_(b,l) = rho [ (a0 phi_b / dt + th um.g_b) d_kl + th phi_b G[l][k] ] - th mu [ lap_b d_kl + H_b[k][l] ]
// emit(slot, value) with slot = (a*6+b)*9 + ri*3 + ci.
// One pass integrates the block forms in MASK (bit 0 J_uu, 1 J_up, 2 J_pu, 3 J_pp) with one rule; test node and mask
// are compile-time constants, so every accumulator index is static (registers, no stack) and only the accumulators of
// the blocks in the pass exist.
#define P2_UU 1
#define P2_UP 2
#define P2_PU 4
#define P2_PP 8

template <int a, int MASK, typename Emit>
HEMO_HD void p2_jac_pass(const P2Cell& c, const HemoForm& par, const HemoP2Rule& ru, Emit emit) {
    const double th = par.theta, rho = par.rho, mu = par.mu;
    double uu[(MASK & P2_UU) ? 6 : 1][2][2], up[(MASK & P2_UP) ? 6 : 1][2], pu[(MASK & P2_PU) ? 6 : 1][2], pp[(MASK & P2_PP) ? 6 : 1];
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        if (MASK & P2_UU) uu[b][0][0] = uu[b][0][1] = uu[b][1][0] = uu[b][1][1] = 0.0;
        if (MASK & P2_UP) up[b][0] = up[b][1] = 0.0;
        if (MASK & P2_PU) pu[b][0] = pu[b][1] = 0.0;
        if (MASK & P2_PP) pp[b] = 0.0;
    }
    for (int q = 0; q < ru.nq; ++q) {
        P2Point s;
        p2_point(c, par, ru.pt[q][0], ru.pt[q][1], s);
        const double w = ru.pt[q][2] * c.adet;
        const double pa = s.phi[a], ga0 = s.g[a][0], ga1 = s.g[a][1], sa = s.umg[a];
        const double ta = pa + s.tau * sa;                 // Galerkin + SUPG weight of the rho part of dR
        const double gak[2] = {ga0, ga1};
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            const double pb = s.phi[b], gb0 = s.g[b][0], gb1 = s.g[b][1];
            const double gbk[2] = {gb0, gb1};
            const double cb = rho * (par.a0_dt * pb + th * s.umg[b]);
            const double hb = th * rho * pb;
            const double vb = th * mu;
            const double Hb[2][2] = {{c.hess[b][0], c.hess[b][1]}, {c.hess[b][1], c.hess[b][2]}};
            if (MASK & P2_UU) {
                const double dd = ga0 * gb0 + ga1 * gb1;
#pragma unroll
                for (int k = 0; k < 2; ++k)
#pragma unroll
                    for (int l = 0; l < 2; ++l) {
                        const double dkl = (k == l) ? 1.0 : 0.0;
                        const double C = cb * dkl + hb * s.G[l][k];                 // rho part of dR_k/dU_bl
                        const double V = vb * (c.lapb[b] * dkl + Hb[k][l]);         // viscous strong part
                        uu[b][k][l] += w * (ta * C - s.tau * sa * V + vb * (dd * dkl + gak[l] * gbk[k]) +
                                            th * s.tau * s.R[k] * pb * gak[l] + th * s.taul * rho * gak[k] * gbk[l]);
                    }
            }
            if (MASK & P2_UP) {
                up[b][0] += w * (-pb * ga0 + s.tau * sa * gb0);
                up[b][1] += w * (-pb * ga1 + s.tau * sa * gb1);
            }
            if (MASK & P2_PU) {
#pragma unroll
                for (int l = 0; l < 2; ++l) {
                    double t = 0.0;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const double dkl = (k == l) ? 1.0 : 0.0;
                        t += (cb * dkl + hb * s.G[l][k] - vb * (c.lapb[b] * dkl + Hb[k][l])) * gak[k];
                    }
                    pu[b][l] += w * (th * pa * gbk[l] + s.tau * par.inv_rho * t);
                }
            }
            if (MASK & P2_PP) pp[b] += w * s.tau * par.inv_rho * (ga0 * gb0 + ga1 * gb1);
        }
    }
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        const int base = (a * 6 + b) * 9;
        if (MASK & P2_UU) {
            emit(base + 0, uu[b][0][0]); emit(base + 1, uu[b][0][1]);
            emit(base + 3, uu[b][1][0]); emit(base + 4, uu[b][1][1]);
        }
        if (MASK & P2_UP) { emit(base + 2, up[b][0]); emit(base + 5, up[b][1]); }
        if (MASK & P2_PU) { emit(base + 6, pu[b][0]); emit(base + 7, pu[b][1]); }
        if (MASK & P2_PP) emit(base + 8, pp[b]);
    }
}

// run-time test node
template <int MASK, typename Emit>
HEMO_HD void p2_jac_pass_node(const P2Cell& c, const HemoForm& par, const HemoP2Rule& ru, int a, Emit emit) {
    switch (a) {
        case 0: p2_jac_pass<0, MASK>(c, par, ru, emit); break;
        case 1: p2_jac_pass<1, MASK>(c, par, ru, emit); break;
        case 2: p2_jac_pass<2, MASK>(c, par, ru, emit); break;
        case 3: p2_jac_pass<3, MASK>(c, par, ru, emit); break;
        case 4: p2_jac_pass<4, MASK>(c, par, ru, emit); break;
        default: p2_jac_pass<5, MASK>(c, par, ru, emit); break;
    }
}

// Passes of the Jacobian: block forms with identical rules (aliases) are integrated together when the combination is
// one of the fused ones (J_up + J_pu; all four), otherwise block by block.  passes[i] = (rule id, mask); returns count.
static inline int hemo_p2_jacobian_passes(const HemoP2Rule* rules, int passes[4][2]) {
    bool done[6] = {true, true, false, false, false, false};
    int np = 0;
    for (int r0 = HEMO_Q_UU; r0 <= HEMO_Q_PP; ++r0) {
        if (done[r0]) continue;
        int mask = 0;
        for (int r = r0; r <= HEMO_Q_PP; ++r)
            if (!done[r] && (r == r0 || rules[r].alias == r0)) mask |= 1 << (r - HEMO_Q_UU);
        if (mask != (P2_UP | P2_PU) && mask != 15) mask = 1 << (r0 - HEMO_Q_UU);      // not a fused combination: r0 alone
        for (int r = HEMO_Q_UU; r <= HEMO_Q_PP; ++r)
            if (mask & (1 << (r - HEMO_Q_UU))) done[r] = true;
        passes[np][0] = r0; passes[np][1] = mask;
        ++np;
    }
    return np;
}

template <typename Emit>
HEMO_HD void p2_jac_pass_rt(const P2Cell& c, const HemoForm& par, const HemoP2Rule& ru, int a, int mask, Emit emit) {
    switch (mask) {
        case P2_UU: p2_jac_pass_node<P2_UU>(c, par, ru, a, emit); break;
        case P2_UP: p2_jac_pass_node<P2_UP>(c, par, ru, a, emit); break;
        case P2_PU: p2_jac_pass_node<P2_PU>(c, par, ru, a, emit); break;
        case P2_PP: p2_jac_pass_node<P2_PP>(c, par, ru, a, emit); break;
        case (P2_UP | P2_PU): p2_jac_pass_node<(P2_UP | P2_PU)>(c, par, ru, a, emit); break;
        default: p2_jac_pass_node<15>(c, par, ru, a, emit); break;
    }
}

// all rows of test node a (host checks, lifting of boundary-adjacent cells)
template <typename Emit>
inline void p2_cell_jacobian_rows(const P2Cell& c, const HemoForm& par, const HemoP2Rule* rules, int a, Emit emit) {
    int passes[4][2];
    const int np = hemo_p2_jacobian_passes(rules, passes);
    for (int i = 0; i < np; ++i) p2_jac_pass_rt(c, par, rules[passes[i][0]], a, passes[i][1], emit);
}

// ---- exterior facets ------------------------------------------------------------------------------------------------
// The boundary terms (src/solvers/stabilized_schur.py:79; stabilized_schur_pressure_backflow.py:192-217) are affine in
// (u_m, p): one routine evaluates them for given nodal values, with or without the constant pressure term; the
// residual uses it once, the Jacobian / the lifting apply it to unit vectors (d u_m / d u = theta).  All six basis
// functions have non-zero gradients on a facet, so every test node gets a row (symmetric Nitsche term).
HEMO_HD void p2_facet_eval(const P2Cell& c, const HemoForm& par, const HemoFacetRule& fr, const hemo_facet_coef& co, int mask,
                           const double Um[6][2], const double Pv[6], bool with_const, double Fu[6][2]) {
#pragma unroll
    for (int a = 0; a < 6; ++a) Fu[a][0] = Fu[a][1] = 0.0;
    const double vref[3][2] = {{0.0, 0.0}, {1.0, 0.0}, {0.0, 1.0}};
    for (int lf = 0; lf < 3; ++lf) {
        if (!(mask & (1 << lf))) continue;
        int va, vb;
        p2_edge_verts(lf, va, vb);
        const double tx = c.X[vb][0] - c.X[va][0], ty = c.X[vb][1] - c.X[va][1];
        const double len = sqrt(tx * tx + ty * ty);
        double nr[2] = {ty / len, -tx / len};
        const double sgn = (nr[0] * (c.X[va][0] - c.X[lf][0]) + nr[1] * (c.X[va][1] - c.X[lf][1])) >= 0.0 ? 1.0 : -1.0;
        nr[0] *= sgn; nr[1] *= sgn;
        const double Pn[2][2] = {{1.0 - nr[0] * nr[0], -nr[0] * nr[1]}, {-nr[1] * nr[0], 1.0 - nr[1] * nr[1]}};
        for (int q = 0; q < fr.nq; ++q) {
            const double sq = fr.s[q], w = fr.w[q] * len;
            const double xi = (1.0 - sq) * vref[va][0] + sq * vref[vb][0], eta = (1.0 - sq) * vref[va][1] + sq * vref[vb][1];
            double phi[6], dr[6][2], g[6][2];
            p2_tabulate(xi, eta, phi, dr);
            double um[2] = {0, 0}, un[2] = {0, 0}, G[2][2] = {{0, 0}, {0, 0}}, p = 0.0;
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                g[a][0] = dr[a][0] * c.K[0][0] + dr[a][1] * c.K[1][0];
                g[a][1] = dr[a][0] * c.K[0][1] + dr[a][1] * c.K[1][1];
                um[0] += phi[a] * Um[a][0]; um[1] += phi[a] * Um[a][1];
                un[0] += phi[a] * c.N[a][0]; un[1] += phi[a] * c.N[a][1];
                p += phi[a] * Pv[a];
                G[0][0] += g[a][0] * Um[a][0]; G[0][1] += g[a][0] * Um[a][1];
                G[1][0] += g[a][1] * Um[a][0]; G[1][1] += g[a][1] * Um[a][1];
            }
            const double e01 = 0.5 * (G[0][1] + G[1][0]);
            const double Gn[2] = {G[0][0] * nr[0] + G[0][1] * nr[1], G[1][0] * nr[0] + G[1][1] * nr[1]};
            const double en[2] = {G[0][0] * nr[0] + e01 * nr[1], e01 * nr[0] + G[1][1] * nr[1]};
            const double umT[2] = {Pn[0][0] * um[0] + Pn[0][1] * um[1], Pn[1][0] * um[0] + Pn[1][1] * um[1]};
            const double enT[2] = {Pn[0][0] * en[0] + Pn[1][0] * en[1], Pn[0][1] * en[0] + Pn[1][1] * en[1]};
            const double unn = un[0] * nr[0] + un[1] * nr[1];
            const double unm = 0.5 * (unn - fabs(unn));
            const double pc = co.a_p * p + (with_const ? co.pconst : 0.0);
            const double pen = co.a_n * co.beta_n * par.mu / c.h;
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                const double dn = g[a][0] * nr[0] + g[a][1] * nr[1];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    double v = pc * phi[a] * nr[k] - co.a_g * par.mu * phi[a] * Gn[k] - co.a_s * 2.0 * par.mu * phi[a] * en[k];
                    if (co.a_n != 0.0) {
                        // eps(v) n for v = phi_a e_k: 1/2 (g_a[i] n_k + (g_a.n) d_ki), contracted with u_T
                        const double ev = 0.5 * ((g[a][0] * umT[0] + g[a][1] * umT[1]) * nr[k] + dn * umT[k]);
                        v += -co.a_n * 2.0 * par.mu * phi[a] * enT[k] - co.a_n * 2.0 * par.mu * ev + pen * phi[a] * umT[k];
                    }
                    if (co.a_b != 0.0) v -= co.a_b * co.beta_b * par.rho * unm * phi[a] * um[k];
                    Fu[a][k] += w * v;
                }
            }
        }
    }
}

// residual contribution of the tagged facets of a cell
HEMO_HD void p2_facet_residual(const P2Cell& c, const HemoForm& par, const HemoFacetRule& fr, const hemo_facet_coef& co, int mask,
                               double Fu[6][2]) {
    double Um[6][2];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        Um[a][0] = par.theta * c.U[a][0] + (1.0 - par.theta) * c.N[a][0];
        Um[a][1] = par.theta * c.U[a][1] + (1.0 - par.theta) * c.N[a][1];
    }
    p2_facet_eval(c, par, fr, co, mask, Um, c.P, true, Fu);
}

// column (b, ci) of the facet Jacobian: ci < 2 velocity component (chain factor theta), ci == 2 pressure
HEMO_HD void p2_facet_column(const P2Cell& c, const HemoForm& par, const HemoFacetRule& fr, const hemo_facet_coef& co, int mask,
                             int b, int ci, double col[6][2]) {
    double Um[6][2], Pv[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) { Um[a][0] = Um[a][1] = 0.0; Pv[a] = 0.0; }
    if (ci < 2) Um[b][ci] = par.theta;
    else Pv[b] = 1.0;
    p2_facet_eval(c, par, fr, co, mask, Um, Pv, false, col);
}

// int u_prev . n over the tagged facets (outlet flux) with the facet rule
HEMO_HD double p2_cell_flux(const P2Cell& c, const HemoFacetRule& fr, int mask) {
    const double vref[3][2] = {{0.0, 0.0}, {1.0, 0.0}, {0.0, 1.0}};
    double total = 0.0;
    for (int lf = 0; lf < 3; ++lf) {
        if (!(mask & (1 << lf))) continue;
        int va, vb;
        p2_edge_verts(lf, va, vb);
        const double tx = c.X[vb][0] - c.X[va][0], ty = c.X[vb][1] - c.X[va][1];
        double nr[2] = {ty, -tx};                                  // |nr| = facet length
        const double sgn = (nr[0] * (c.X[va][0] - c.X[lf][0]) + nr[1] * (c.X[va][1] - c.X[lf][1])) >= 0.0 ? 1.0 : -1.0;
        for (int q = 0; q < fr.nq; ++q) {
            const double sq = fr.s[q];
            const double xi = (1.0 - sq) * vref[va][0] + sq * vref[vb][0], eta = (1.0 - sq) * vref[va][1] + sq * vref[vb][1];
            double phi[6], dr[6][2];
            p2_tabulate(xi, eta, phi, dr);
            double un[2] = {0, 0};
#pragma unroll
            for (int a = 0; a < 6; ++a) { un[0] += phi[a] * c.N[a][0]; un[1] += phi[a] * c.N[a][1]; }
            total += fr.w[q] * sgn * (un[0] * nr[0] + un[1] * nr[1]);
        }
    }
    return total;
}

// P2 stiffness matrix (exact: 3-point edge-midpoint rule) and HRZ-lumped mass (diagonal of the consistent mass scaled to
// the cell area) for the Schur-complement approximation
HEMO_HD void p2_cell_laplace_mass(const P2Cell& c, double Ke[6][6], double Me[6]) {
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = 0; b < 6; ++b) Ke[a][b] = 0.0;
    const double qp[3][2] = {{0.5, 0.5}, {0.0, 0.5}, {0.5, 0.0}};
    for (int q = 0; q < 3; ++q) {
        double phi[6], dr[6][2], g[6][2];
        p2_tabulate(qp[q][0], qp[q][1], phi, dr);
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            g[a][0] = dr[a][0] * c.K[0][0] + dr[a][1] * c.K[1][0];
            g[a][1] = dr[a][0] * c.K[0][1] + dr[a][1] * c.K[1][1];
        }
        const double w = c.adet / 6.0;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b < 6; ++b) Ke[a][b] += w * (g[a][0] * g[b][0] + g[a][1] * g[b][1]);
    }
    const double area = 0.5 * c.adet;
    const double dv = 1.0 / 30.0, de = 8.0 / 45.0, tot = 3.0 * dv + 3.0 * de;
#pragma unroll
    for (int a = 0; a < 3; ++a) { Me[a] = area * dv / tot; Me[3 + a] = area * de / tot; }
}
