// Element routines of the stabilized Navier–Stokes forms on Q1–Q1 quadrilaterals
// (tensor-ordered vertices (0,0),(1,0),(0,1),(1,1); bilinear, i.e. non-affine, geometry).
//
// Replaces the FFCx `tabulate_tensor` kernels DOLFINx runs for the forms of
// src/solvers/stabilized_schur.py:60-123 when `mesh.topology.cell_name()` is
// "quadrilateral" (reference src/scenarios/stenosis_pressure_structured.py:379-386:
// transfinite + recombined mesh).  The full integrand is evaluated at every point of
// each block form's own rule (:188-189).  Because the map is bilinear the physical
// Hessian of a basis function is  H(phi_a) = theta_a * kappa  with
//     kappa_ij = K_0i K_1j + K_1i K_0j,   theta_a = s_a - grad(phi_a) . (X0 - X1 - X2 + X3),
// which is what `div(sigma(u_mid, p))` in the strong residual (:95-97) needs.
//
// The routines are `__host__ __device__`: the CUDA kernels in assembly_q1.cu are thin
// load / call / store wrappers, and tests/host_q1 compiles this same header with g++ to
// check the arithmetic against the numpy oracle on machines without a GPU (test
// infrastructure only; the product path never runs it on the CPU).
#pragma once
#include <math.h>

#include "../../include/hemo.h"
#include "hemo_rules.h"

#ifdef __CUDACC__
#define HEMO_HD __host__ __device__ __forceinline__
#else
#define HEMO_HD inline
#endif

// block bits
#define Q1_UU 1
#define Q1_UP 2
#define Q1_PU 4
#define Q1_PP 8

struct Q1Cell {
    double X[4][2];
    double U[4][2], N[4][2], P[4];
    double H[4][2];     // history of the time derivative: dudt = (a0 U - H) / dt (mid-point scheme: H = N)
    double h;
    // per-cell constants of tau_supg / tau_lsic (q1_prepare): hoisted out of the quadrature loops
    double inv_h2, c23, re_fac, half_h;
};

struct Q1Geom {
    double phi[4], g[4][2], theta[4];
    double k[2][2], trk;
    double adet;
};

struct Q1State {
    double um[2], G[2][2], R[2], acc[2], divu, p, tau, taul, umg[4];
};

HEMO_HD void q1_geom(const Q1Cell& c, double xi, double eta, Q1Geom& o) {
    const double xm = 1.0 - xi, em = 1.0 - eta;
    o.phi[0] = xm * em; o.phi[1] = xi * em; o.phi[2] = xm * eta; o.phi[3] = xi * eta;
    const double dr[4][2] = {{-em, -xm}, {em, -xi}, {-eta, xm}, {eta, xi}};
    double J[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
            J[i][j] = c.X[0][i] * dr[0][j] + c.X[1][i] * dr[1][j] + c.X[2][i] * dr[2][j] + c.X[3][i] * dr[3][j];
    const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double id = 1.0 / det;
    const double K[2][2] = {{J[1][1] * id, -J[0][1] * id}, {-J[1][0] * id, J[0][0] * id}};
    const double cx[2] = {c.X[0][0] - c.X[1][0] - c.X[2][0] + c.X[3][0],
                          c.X[0][1] - c.X[1][1] - c.X[2][1] + c.X[3][1]};
    const double sg[4] = {1.0, -1.0, -1.0, 1.0};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        o.g[a][0] = dr[a][0] * K[0][0] + dr[a][1] * K[1][0];
        o.g[a][1] = dr[a][0] * K[0][1] + dr[a][1] * K[1][1];
        o.theta[a] = sg[a] - (o.g[a][0] * cx[0] + o.g[a][1] * cx[1]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) o.k[i][j] = K[0][i] * K[1][j] + K[1][i] * K[0][j];
    o.trk = o.k[0][0] + o.k[1][1];
    o.adet = fabs(det);
}

// Per-cell constants of the stabilization parameters; call once after the cell data is loaded.
HEMO_HD void q1_prepare(Q1Cell& c, const HemoForm& par) {
    const double h = c.h;
    c.inv_h2 = 1.0 / (h * h);
    const double t2inv = 2.0 * par.inv_dt;            // 1 / tau_supg2
    const double t3inv = 4.0 * par.nu * c.inv_h2;     // 1 / tau_supg3
    c.c23 = t2inv * t2inv + t3inv * t3inv;
    c.re_fac = h / (2.0 * par.nu);                    // Re = |u_n| h / (2 nu)
    c.half_h = 0.5 * h;
}

// tau_supg and tau_lsic at a point (stabilized_schur.py:91-118)
HEMO_HD void q1_tau(const HemoForm& par, const Q1Cell& c, double unx, double uny, double& tau, double& taul) {
    const double v2 = unx * unx + uny * uny;
    const double t1 = fmax(4.0 * v2, par.eps0 * par.eps0) * c.inv_h2;   // (max(2|u_n|, eps)/h)^2
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

HEMO_HD void q1_state(const Q1Cell& c, const HemoForm& par, const Q1Geom& ge, Q1State& s) {
    double u[2] = {0, 0}, un[2] = {0, 0}, uh[2] = {0, 0}, wv[2] = {0, 0}, gp[2] = {0, 0};
    const double th = par.theta;
    s.p = 0.0;
    s.G[0][0] = s.G[0][1] = s.G[1][0] = s.G[1][1] = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const double m0 = th * c.U[a][0] + (1.0 - th) * c.N[a][0], m1 = th * c.U[a][1] + (1.0 - th) * c.N[a][1];
        u[0] += ge.phi[a] * c.U[a][0]; u[1] += ge.phi[a] * c.U[a][1];
        un[0] += ge.phi[a] * c.N[a][0]; un[1] += ge.phi[a] * c.N[a][1];
        uh[0] += ge.phi[a] * c.H[a][0]; uh[1] += ge.phi[a] * c.H[a][1];
        s.p += ge.phi[a] * c.P[a];
        s.G[0][0] += ge.g[a][0] * m0; s.G[0][1] += ge.g[a][0] * m1;
        s.G[1][0] += ge.g[a][1] * m0; s.G[1][1] += ge.g[a][1] * m1;
        gp[0] += ge.g[a][0] * c.P[a]; gp[1] += ge.g[a][1] * c.P[a];
        wv[0] += ge.theta[a] * m0; wv[1] += ge.theta[a] * m1;
    }
    s.um[0] = th * u[0] + (1.0 - th) * un[0]; s.um[1] = th * u[1] + (1.0 - th) * un[1];
    s.divu = s.G[0][0] + s.G[1][1];
    const double idt = par.inv_dt;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const double conv = s.um[0] * s.G[0][k] + s.um[1] * s.G[1][k];
        // div(2 mu eps(u_m)) = mu (lap u_m + grad div u_m)
        const double visc = par.mu * (wv[k] * ge.trk + ge.k[k][0] * wv[0] + ge.k[k][1] * wv[1]);
        s.acc[k] = (par.a0 * u[k] - uh[k]) * idt + conv - par.f[k];
        s.R[k] = par.rho * s.acc[k] + gp[k] - visc;
    }
    q1_tau(par, c, un[0], un[1], s.tau, s.taul);
#pragma unroll
    for (int a = 0; a < 4; ++a) s.umg[a] = s.um[0] * ge.g[a][0] + s.um[1] * ge.g[a][1];
}

// ---- residual ---------------------------------------------------------------------
// One quadrature point of F_u (do_u) and / or F_p (do_p); w = weight * |det J|.
HEMO_HD void q1_residual_point(const HemoForm& par, const Q1Geom& ge, const Q1State& s, double w,
                               bool do_u, bool do_p, double Fu[4][2], double Fp[4]) {
    const double rho = par.rho, mu = par.mu;
    if (do_u) {
        const double e01 = 0.5 * (s.G[0][1] + s.G[1][0]);
        const double sig[2][2] = {{2.0 * mu * s.G[0][0] - s.p, 2.0 * mu * e01},
                                  {2.0 * mu * e01, 2.0 * mu * s.G[1][1] - s.p}};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int k = 0; k < 2; ++k)
                Fu[a][k] += w * (rho * ge.phi[a] * s.acc[k] + ge.g[a][0] * sig[0][k] + ge.g[a][1] * sig[1][k] +
                                 s.tau * s.umg[a] * s.R[k] + s.taul * rho * s.divu * ge.g[a][k]);
    }
    if (do_p) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
            Fp[a] += w * (ge.phi[a] * s.divu + s.tau * par.inv_rho * (s.R[0] * ge.g[a][0] + s.R[1] * ge.g[a][1]));
    }
}

// Element residual with the rules of the F_u and F_p block forms (ids HEMO_Q_FU, HEMO_Q_FP).
HEMO_HD void q1_cell_residual(const Q1Cell& c, const HemoForm& par, const HemoQuadRule* rules,
                              double Fu[4][2], double Fp[4]) {
#pragma unroll
    for (int a = 0; a < 4; ++a) { Fu[a][0] = Fu[a][1] = 0.0; Fp[a] = 0.0; }
    const bool shared = rules[HEMO_Q_FP].alias == HEMO_Q_FU;
    for (int r = HEMO_Q_FU; r <= HEMO_Q_FP; ++r) {
        if (r == HEMO_Q_FP && shared) break;
        const HemoQuadRule& ru = rules[r];
        const bool do_u = (r == HEMO_Q_FU), do_p = (r == HEMO_Q_FP) || shared;
        for (int q = 0; q < ru.nq; ++q) {
            Q1Geom ge;
            Q1State s;
            q1_geom(c, ru.pt[q][0], ru.pt[q][1], ge);
            q1_state(c, par, ge, s);
            q1_residual_point(par, ge, s, ru.pt[q][2] * ge.adet, do_u, do_p, Fu, Fp);
        }
    }
}

// ---- Jacobian ---------------------------------------------------------------------
// The 12 x 12 element tensor is produced by three work items per cell so that the point set-up
// (geometry, state, tau: one division, two square roots and ~200 flops) is repeated as little as
// the register file allows:
//   J_uu rows of test nodes {0,1} and {2,3}: 2 x 32 accumulators, rule of the J_uu form;
//   J_up + J_pu (64 accumulators, their rule(s)) followed by J_pp (16 accumulators, its rule).
// Slots follow the element buffer layout: (a*4+b)*9 + ri*3 + ci with ri/ci in (u_x, u_y, p).

// One quadrature point of J_uu for the test nodes A0, A0+1: uu[i][b][k*2+l].
// With C_b[k][l] = cb d_kl + hb G[l][k] and V_b[k][l] = vb (tr(kappa) d_kl + kappa[k][l])
// (dR = C - V) the integrand
//   phi_a C + th mu (g_a.g_b d_kl + g_a[l] g_b[k]) + tau (s_a (C - V) + th phi_b g_a[l] R[k]) + th tau_l rho g_a[k] g_b[l]
// is regrouped into per-test-node factors times per-trial-node factors, 5 fused multiply-adds per entry:
//   (phi_a + tau s_a) C - tau s_a V + (th mu g_a[l]) g_b[k] + (th tau g_a[l] R[k]) phi_b + (th tau_l rho g_a[k]) g_b[l]
//   + th mu g_a.g_b d_kl
template <int A0>
HEMO_HD void q1_uu_point(const HemoForm& par, const Q1Geom& ge, const Q1State& s, double w, double uu[2][4][4]) {
    // th = d(u_e)/du and idt = d(dudt)/du carry the time scheme (1/2 and 1/dt for the mid-point rule)
    const double rho = par.rho, mu = par.mu, idt = par.a0_dt, th = par.theta;
    double a1[2], a2[2], a3[2][2], a4[2][2][2], a5[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int a = A0 + i;
        a2[i] = w * s.tau * s.umg[a];
        a1[i] = w * ge.phi[a] + a2[i];
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            a3[i][l] = w * th * mu * ge.g[a][l];
            a5[i][l] = w * th * s.taul * rho * ge.g[a][l];
            const double t = w * th * s.tau * ge.g[a][l];
            a4[i][l][0] = t * s.R[0];
            a4[i][l][1] = t * s.R[1];
        }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const double cb = rho * (ge.phi[b] * idt + th * s.umg[b]);
        const double hb = th * rho * ge.phi[b];
        const double vb = th * mu * ge.theta[b];
        double C[2][2], V[2][2];
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int l = 0; l < 2; ++l) {
                const double dkl = (k == l) ? 1.0 : 0.0;
                C[k][l] = cb * dkl + hb * s.G[l][k];
                V[k][l] = vb * (ge.trk * dkl + ge.k[k][l]);
            }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            // th mu g_a.g_b on the diagonal (k == l)
            const double dg = a3[i][0] * ge.g[b][0] + a3[i][1] * ge.g[b][1];
#pragma unroll
            for (int k = 0; k < 2; ++k)
#pragma unroll
                for (int l = 0; l < 2; ++l) {
                    double v = a1[i] * C[k][l] - a2[i] * V[k][l];
                    v += a3[i][l] * ge.g[b][k];
                    v += a4[i][l][k] * ge.phi[b];
                    v += a5[i][k] * ge.g[b][l];
                    if (k == l) v += dg;
                    uu[i][b][k * 2 + l] += v;
                }
        }
    }
}

// One quadrature point of J_up (up[a][b][k]) and / or J_pu (pu[a][b][l]).
HEMO_HD void q1_uppu_point(const HemoForm& par, const Q1Geom& ge, const Q1State& s, double w, bool do_up,
                           bool do_pu, double up[4][4][2], double pu[4][4][2]) {
    const double rho = par.rho, mu = par.mu, idt = par.a0_dt, th = par.theta;
    const double tr = s.tau * par.inv_rho;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        double dR[2][2];
        if (do_pu) {
            const double cb = rho * (ge.phi[b] * idt + th * s.umg[b]);
            const double hb = th * rho * ge.phi[b];
            const double vb = th * mu * ge.theta[b];
#pragma unroll
            for (int k = 0; k < 2; ++k)
#pragma unroll
                for (int l = 0; l < 2; ++l) {
                    const double dkl = (k == l) ? 1.0 : 0.0;
                    dR[k][l] = cb * dkl + hb * s.G[l][k] - vb * (ge.trk * dkl + ge.k[k][l]);
                }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            if (do_up) {
                up[a][b][0] += w * (-ge.phi[b] * ge.g[a][0] + s.tau * s.umg[a] * ge.g[b][0]);
                up[a][b][1] += w * (-ge.phi[b] * ge.g[a][1] + s.tau * s.umg[a] * ge.g[b][1]);
            }
            if (do_pu) {
                pu[a][b][0] += w * (th * ge.phi[a] * ge.g[b][0] + tr * (dR[0][0] * ge.g[a][0] + dR[1][0] * ge.g[a][1]));
                pu[a][b][1] += w * (th * ge.phi[a] * ge.g[b][1] + tr * (dR[0][1] * ge.g[a][0] + dR[1][1] * ge.g[a][1]));
            }
        }
    }
}

// Blocks integrated with rule r (block ids HEMO_Q_UU..HEMO_Q_PP): those whose alias is r.
HEMO_HD int q1_blocks_of_rule(const HemoQuadRule* rules, int r) {
    int m = 0;
    if (rules[HEMO_Q_UU].alias == r) m |= Q1_UU;
    if (rules[HEMO_Q_UP].alias == r) m |= Q1_UP;
    if (rules[HEMO_Q_PU].alias == r) m |= Q1_PU;
    if (rules[HEMO_Q_PP].alias == r) m |= Q1_PP;
    return m;
}

// Work items 0 / 1: J_uu rows of test nodes A0, A0+1 with the J_uu rule; emit(slot, value).
template <int A0, typename Emit>
HEMO_HD void q1_cell_jacobian_uu(const Q1Cell& c, const HemoForm& par, const HemoQuadRule* rules, Emit emit) {
    double uu[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int k = 0; k < 4; ++k) uu[i][b][k] = 0.0;
    const HemoQuadRule& ru = rules[HEMO_Q_UU];
    for (int q = 0; q < ru.nq; ++q) {
        Q1Geom ge;
        Q1State s;
        q1_geom(c, ru.pt[q][0], ru.pt[q][1], ge);
        q1_state(c, par, ge, s);
        q1_uu_point<A0>(par, ge, s, ru.pt[q][2] * ge.adet, uu);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int k = 0; k < 2; ++k)
#pragma unroll
                for (int l = 0; l < 2; ++l) emit(((A0 + i) * 4 + b) * 9 + k * 3 + l, uu[i][b][k * 2 + l]);
}

// Work item 2: J_up and J_pu (one pass when their rules coincide), then J_pp.
template <typename Emit>
HEMO_HD void q1_cell_jacobian_p(const Q1Cell& c, const HemoForm& par, const HemoQuadRule* rules, Emit emit) {
    {
        double up[4][4][2], pu[4][4][2];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) up[a][b][0] = up[a][b][1] = pu[a][b][0] = pu[a][b][1] = 0.0;
        const bool shared = rules[HEMO_Q_PU].alias == rules[HEMO_Q_UP].alias;
        for (int pass = 0; pass < (shared ? 1 : 2); ++pass) {
            const HemoQuadRule& ru = rules[pass == 0 ? HEMO_Q_UP : HEMO_Q_PU];
            const bool do_up = pass == 0, do_pu = shared || pass == 1;
            for (int q = 0; q < ru.nq; ++q) {
                Q1Geom ge;
                Q1State s;
                q1_geom(c, ru.pt[q][0], ru.pt[q][1], ge);
                q1_state(c, par, ge, s);
                q1_uppu_point(par, ge, s, ru.pt[q][2] * ge.adet, do_up, do_pu, up, pu);
            }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                emit((a * 4 + b) * 9 + 2, up[a][b][0]);
                emit((a * 4 + b) * 9 + 5, up[a][b][1]);
                emit((a * 4 + b) * 9 + 6, pu[a][b][0]);
                emit((a * 4 + b) * 9 + 7, pu[a][b][1]);
            }
    }
    double pp[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) pp[a][b] = 0.0;
    const HemoQuadRule& ru = rules[HEMO_Q_PP];
    for (int q = 0; q < ru.nq; ++q) {
        Q1Geom ge;
        q1_geom(c, ru.pt[q][0], ru.pt[q][1], ge);
        // only tau is needed from the state: previous velocity at the point
        double un0 = 0.0, un1 = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) { un0 += ge.phi[a] * c.N[a][0]; un1 += ge.phi[a] * c.N[a][1]; }
        double tau, taul;
        q1_tau(par, c, un0, un1, tau, taul);
        const double wt = ru.pt[q][2] * ge.adet * tau * par.inv_rho;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) pp[a][b] += wt * (ge.g[a][0] * ge.g[b][0] + ge.g[a][1] * ge.g[b][1]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) emit((a * 4 + b) * 9 + 8, pp[a][b]);
}

// ---- lifting ----------------------------------------------------------------------
// F += A_e d with d = (g - x) on constrained dofs (3P apply_lifting with x0 = x, alpha = -1;
// trigger src/solvers/stabilized_schur.py:172-174): the element Jacobian is contracted with
// dl[b] = (dU_x, dU_y, dP) point by point, block by block (each block with its own rule).
HEMO_HD void q1_cell_lift(const Q1Cell& c, const HemoForm& par, const HemoQuadRule* rules,
                          const double dl[4][3], double Fu[4][2], double Fp[4]) {
    const double rho = par.rho, mu = par.mu, idt = par.a0_dt, th = par.theta;
    for (int r = HEMO_Q_UU; r <= HEMO_Q_PP; ++r) {
        const int blocks = q1_blocks_of_rule(rules, r);
        if (blocks == 0) continue;
        const HemoQuadRule& ru = rules[r];
        for (int q = 0; q < ru.nq; ++q) {
            Q1Geom ge;
            Q1State s;
            q1_geom(c, ru.pt[q][0], ru.pt[q][1], ge);
            q1_state(c, par, ge, s);
            const double w = ru.pt[q][2] * ge.adet;
            double du[2] = {0, 0}, dgu[2][2] = {{0, 0}, {0, 0}}, dw[2] = {0, 0}, dp = 0.0, dgp[2] = {0, 0};
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                du[0] += ge.phi[b] * dl[b][0]; du[1] += ge.phi[b] * dl[b][1];
                dgu[0][0] += ge.g[b][0] * dl[b][0]; dgu[0][1] += ge.g[b][0] * dl[b][1];
                dgu[1][0] += ge.g[b][1] * dl[b][0]; dgu[1][1] += ge.g[b][1] * dl[b][1];
                dw[0] += ge.theta[b] * dl[b][0]; dw[1] += ge.theta[b] * dl[b][1];
                dp += ge.phi[b] * dl[b][2];
                dgp[0] += ge.g[b][0] * dl[b][2]; dgp[1] += ge.g[b][1] * dl[b][2];
            }
            double dC[2], dRu[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                dC[k] = rho * (du[k] * idt + th * (du[0] * s.G[0][k] + du[1] * s.G[1][k]) +
                               th * (s.um[0] * dgu[0][k] + s.um[1] * dgu[1][k]));
                dRu[k] = dC[k] - th * mu * (dw[k] * ge.trk + ge.k[k][0] * dw[0] + ge.k[k][1] * dw[1]);
            }
            const double ddiv = dgu[0][0] + dgu[1][1];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const double dug = du[0] * ge.g[a][0] + du[1] * ge.g[a][1];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    double v = 0.0;
                    if (blocks & Q1_UU)
                        v += ge.phi[a] * dC[k] +
                             th * mu * (ge.g[a][0] * (dgu[0][k] + dgu[k][0]) + ge.g[a][1] * (dgu[1][k] + dgu[k][1])) +
                             s.tau * (s.umg[a] * dRu[k] + th * dug * s.R[k]) + th * s.taul * rho * ge.g[a][k] * ddiv;
                    if (blocks & Q1_UP) v += -dp * ge.g[a][k] + s.tau * s.umg[a] * dgp[k];
                    Fu[a][k] += w * v;
                }
                double vp = 0.0;
                if (blocks & Q1_PU)
                    vp += th * ge.phi[a] * ddiv + s.tau * par.inv_rho * (dRu[0] * ge.g[a][0] + dRu[1] * ge.g[a][1]);
                if (blocks & Q1_PP) vp += s.tau * par.inv_rho * (ge.g[a][0] * dgp[0] + ge.g[a][1] * dgp[1]);
                Fp[a] += w * vp;
            }
        }
    }
}

// ---- exterior facets --------------------------------------------------------------
// local facets (0,1),(0,2),(1,3),(2,3) (3P Basix numbering); parameter s runs from the
// first to the second vertex.
HEMO_HD void q1_facet_verts(int lf, int& va, int& vb) {
    va = (lf == 0 || lf == 1) ? 0 : (lf == 2 ? 1 : 2);
    vb = (lf == 0) ? 1 : (lf == 1 ? 2 : 3);
}

HEMO_HD void q1_facet_ref(int lf, double s, double& xi, double& eta) {
    xi = (lf == 0 || lf == 3) ? s : (lf == 1 ? 0.0 : 1.0);
    eta = (lf == 0) ? 0.0 : (lf == 3 ? 1.0 : s);
}

// unit outward normal (away from the cell centroid) and length of a straight edge
HEMO_HD void q1_facet_normal(const Q1Cell& c, int lf, double nr[2], double& len) {
    int va, vb;
    q1_facet_verts(lf, va, vb);
    const double tx = c.X[vb][0] - c.X[va][0], ty = c.X[vb][1] - c.X[va][1];
    len = sqrt(tx * tx + ty * ty);
    double nx = ty / len, ny = -tx / len;
    const double cx = 0.25 * (c.X[0][0] + c.X[1][0] + c.X[2][0] + c.X[3][0]);
    const double cy = 0.25 * (c.X[0][1] + c.X[1][1] + c.X[2][1] + c.X[3][1]);
    const double side = nx * (0.5 * (c.X[va][0] + c.X[vb][0]) - cx) + ny * (0.5 * (c.X[va][1] + c.X[vb][1]) - cy);
    if (side < 0.0) { nx = -nx; ny = -ny; }
    nr[0] = nx; nr[1] = ny;
}

// Facet terms of one boundary cell (see FacetSet in oracle/ns_oracle.py for the term each
// coefficient multiplies; src/solvers/stabilized_schur.py:79,
// src/solvers/stabilized_schur_pressure_backflow.py:192-217).
//   residual(a, k, value)        : += into Fu[a][k]           (want_res)
//   jac(a, b, ri, ci, value)     : += into d Fu[a][ri] / d (U_b,ci | P_b for ci = 2)   (want_jac)
template <typename Res, typename Jac>
HEMO_HD void q1_cell_facets(const Q1Cell& c, const HemoForm& par, const HemoFacetRule& fr,
                            const hemo_facet_coef& co, int mask, bool want_res, bool want_jac, Res residual, Jac jac) {
    const double mu = par.mu, rho = par.rho, th = par.theta;
    const double pen = co.a_n * co.beta_n * mu / c.h;
    const double bf = co.a_b * co.beta_b * rho;
    for (int lf = 0; lf < 4; ++lf) {
        if (!(mask & (1 << lf))) continue;
        double nr[2], len;
        q1_facet_normal(c, lf, nr, len);
        const double Pn[2][2] = {{1.0 - nr[0] * nr[0], -nr[0] * nr[1]}, {-nr[0] * nr[1], 1.0 - nr[1] * nr[1]}};
        for (int q = 0; q < fr.nq; ++q) {
            double xi, eta;
            q1_facet_ref(lf, fr.s[q], xi, eta);
            Q1Geom ge;
            q1_geom(c, xi, eta, ge);
            const double w = fr.w[q] * len;
            double um[2] = {0, 0}, un[2] = {0, 0}, p = 0.0, G[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const double m0 = th * c.U[a][0] + (1.0 - th) * c.N[a][0], m1 = th * c.U[a][1] + (1.0 - th) * c.N[a][1];
                um[0] += ge.phi[a] * m0; um[1] += ge.phi[a] * m1;
                un[0] += ge.phi[a] * c.N[a][0]; un[1] += ge.phi[a] * c.N[a][1];
                p += ge.phi[a] * c.P[a];
                G[0][0] += ge.g[a][0] * m0; G[0][1] += ge.g[a][0] * m1;
                G[1][0] += ge.g[a][1] * m0; G[1][1] += ge.g[a][1] * m1;
            }
            const double unn = un[0] * nr[0] + un[1] * nr[1];
            const double unm = 0.5 * (unn - fabs(unn));
            double dn[4], Png[4][2];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                dn[a] = ge.g[a][0] * nr[0] + ge.g[a][1] * nr[1];
                Png[a][0] = Pn[0][0] * ge.g[a][0] + Pn[0][1] * ge.g[a][1];
                Png[a][1] = Pn[1][0] * ge.g[a][0] + Pn[1][1] * ge.g[a][1];
            }
            if (want_res) {
                const double Gn[2] = {G[0][0] * nr[0] + G[0][1] * nr[1], G[1][0] * nr[0] + G[1][1] * nr[1]};
                const double e01 = 0.5 * (G[0][1] + G[1][0]);
                const double en[2] = {G[0][0] * nr[0] + e01 * nr[1], e01 * nr[0] + G[1][1] * nr[1]};
                const double enT[2] = {Pn[0][0] * en[0] + Pn[0][1] * en[1], Pn[1][0] * en[0] + Pn[1][1] * en[1]};
                const double umT[2] = {Pn[0][0] * um[0] + Pn[0][1] * um[1], Pn[1][0] * um[0] + Pn[1][1] * um[1]};
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const double gT = ge.g[a][0] * umT[0] + ge.g[a][1] * umT[1];
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        double vv = (co.a_p * p + co.pconst) * ge.phi[a] * nr[k];
                        vv -= co.a_g * mu * ge.phi[a] * Gn[k];
                        vv -= co.a_s * 2.0 * mu * ge.phi[a] * en[k];
                        vv -= co.a_n * 2.0 * mu * ge.phi[a] * enT[k];
                        // -(2 mu eps(v) n).u_T with eps(v) n = 1/2 (g_a n_k + (g_a.n) e_k)
                        vv -= co.a_n * mu * (gT * nr[k] + dn[a] * umT[k]);
                        vv += pen * ge.phi[a] * umT[k];
                        vv -= bf * unm * ge.phi[a] * um[k];
                        residual(a, k, w * vv);
                    }
                }
            }
            if (want_jac) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const double pab = ge.phi[a] * ge.phi[b];
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
#pragma unroll
                            for (int l = 0; l < 2; ++l) {
                                const double dkl = (k == l) ? 1.0 : 0.0;
                                double vv = -th * co.a_g * mu * ge.g[b][k] * nr[l] * ge.phi[a];
                                vv -= th * co.a_s * mu * (ge.g[b][k] * nr[l] + dn[b] * dkl) * ge.phi[a];
                                vv -= th * co.a_n * mu * (Png[b][k] * nr[l] + dn[b] * Pn[k][l]) * ge.phi[a];
                                vv -= th * co.a_n * mu * (Png[a][l] * nr[k] + dn[a] * Pn[k][l]) * ge.phi[b];
                                vv += th * pen * Pn[k][l] * pab;
                                vv -= th * bf * unm * pab * dkl;
                                jac(a, b, k, l, w * vv);
                            }
                            jac(a, b, k, 2, w * co.a_p * nr[k] * pab);
                        }
                    }
            }
        }
    }
}

// int u_prev . n ds over the masked facets of one cell (linear trace: mid-point rule is exact)
HEMO_HD double q1_cell_flux(const Q1Cell& c, int mask) {
    double qsum = 0.0;
    for (int lf = 0; lf < 4; ++lf) {
        if (!(mask & (1 << lf))) continue;
        double nr[2], len;
        q1_facet_normal(c, lf, nr, len);
        int va, vb;
        q1_facet_verts(lf, va, vb);
        qsum += 0.5 * len * ((c.N[va][0] + c.N[vb][0]) * nr[0] + (c.N[va][1] + c.N[vb][1]) * nr[1]);
    }
    return qsum;
}

// Pressure Laplacian K_ab = int grad phi_a . grad phi_b and lumped mass M_a = int phi_a with a
// 3 x 3 Gauss rule (operators of the Schur-complement approximation; preconditioner only).
HEMO_HD void q1_cell_laplace_mass(const Q1Cell& c, double Ke[4][4], double Me[4]) {
    const double gp[3] = {0.5 - 0.5 * 0.7745966692414834, 0.5, 0.5 + 0.5 * 0.7745966692414834};
    const double gw[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        Me[a] = 0.0;
#pragma unroll
        for (int b = 0; b < 4; ++b) Ke[a][b] = 0.0;
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            Q1Geom ge;
            q1_geom(c, gp[i], gp[j], ge);
            const double w = gw[i] * gw[j] * ge.adet;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                Me[a] += w * ge.phi[a];
#pragma unroll
                for (int b = 0; b < 4; ++b) Ke[a][b] += w * (ge.g[a][0] * ge.g[b][0] + ge.g[a][1] * ge.g[b][1]);
            }
        }
}
