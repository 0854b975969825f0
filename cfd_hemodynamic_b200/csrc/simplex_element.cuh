// Dimension-generic P1–P1 simplex element routines (D = 2 triangles, D = 3 tetrahedra) of the
// stabilized Navier–Stokes forms, src/solvers/stabilized_schur.py:60-123.
//
// Written for the tetrahedral kernels the north star names (the reference reaches them through
// `mesh.topology.cell_name()`, e.g. src/scenarios/taylor_green.py:34): the arithmetic below is the
// moment factorisation of assembly.cu written once for any D — on an affine simplex every field of
// the integrand is linear in the barycentric coordinates, so the quadrature only enters through
//     T2_ab = |J| sum_q w_q tau(q) phi_a phi_b ,   L0 = |J| sum_q w_q tau_lsic(q)
// of each block form's own rule and the (D+1)^2 node blocks follow in closed form.  It is checked
// on the host against oracle/simplex_oracle.py for D = 2 and D = 3 (tests/test_simplex_host.py,
// compiled with g++: test infrastructure only).  The tetrahedron kernels of assembly_tet.cu
// instantiate it for D = 3.  `simplex_facet` states the exterior-facet terms
// (stabilized_schur.py:79, stabilized_schur_pressure_backflow.py:189-217) once for any D.
#pragma once
#include <math.h>

#include "../../include/hemo.h"
#include "hemo_rules.h"

#ifdef __CUDACC__
#define HEMO_HD __host__ __device__ __forceinline__
#else
#ifndef HEMO_HD
#define HEMO_HD inline
#endif
#endif

#define HEMO_SIMPLEX_MAXQ 343        // collapsed Gauss-Jacobi, 7^3 points (degree 12 on tetrahedra)

template <int D>
struct SimplexRule {
    static constexpr int NV = D + 1;
    int nq;
    double phi[HEMO_SIMPLEX_MAXQ][NV];
    double w[HEMO_SIMPLEX_MAXQ];
    double m0, m1[NV], m2[NV][NV];     // polynomial moments of the rule on the reference cell
};

template <int D>
inline void simplex_rule_set(SimplexRule<D>& r, const double* pts, const double* wts, int nq) {
    constexpr int NV = D + 1;
    r.nq = nq;
    r.m0 = 0.0;
    for (int a = 0; a < NV; ++a) { r.m1[a] = 0.0; for (int b = 0; b < NV; ++b) r.m2[a][b] = 0.0; }
    for (int q = 0; q < nq; ++q) {
        double s = 0.0;
        for (int i = 0; i < D; ++i) { r.phi[q][i + 1] = pts[D * q + i]; s += pts[D * q + i]; }
        r.phi[q][0] = 1.0 - s;
        r.w[q] = wts[q];
        r.m0 += wts[q];
        for (int a = 0; a < NV; ++a) {
            r.m1[a] += wts[q] * r.phi[q][a];
            for (int b = 0; b < NV; ++b) r.m2[a][b] += wts[q] * r.phi[q][a] * r.phi[q][b];
        }
    }
}

template <int D>
struct SimplexCell {
    static constexpr int NV = D + 1;
    double g[NV][D];                   // grad phi_a
    double detJ;                       // |det J|
    double U[NV][D], N[NV][D], H[NV][D], P[NV];
    double h;
    double fbody[D];                   // body force f (hemo_params carries two components; 3-D adds the third)
    // derived (simplex_derive)
    double M[NV][D];                   // u_e = theta U + (1 - theta) N
    double G[D][D];                    // G_ij = d_i u_ej
    double gp[D];                      // grad p
    double s[NV][NV];                  // s[c][a] = M_c . g_a
    double A[NV][D];                   // nodal (a0 u - u_h)/dt + (u_e.grad) u_e - f
    double R[NV][D];                   // nodal strong residual rho A + grad p
    double divu;
};

// geometry from the vertex coordinates X[a][i]
template <int D>
HEMO_HD void simplex_geometry(SimplexCell<D>& c, const double X[D + 1][D]) {
    double J[D][D], K[D][D];
    #pragma unroll
    for (int i = 0; i < D; ++i)
        #pragma unroll
        for (int j = 0; j < D; ++j) J[i][j] = X[j + 1][i] - X[0][i];
    double det;
    if constexpr (D == 2) {
        det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        K[0][0] = J[1][1] / det; K[0][1] = -J[0][1] / det;
        K[1][0] = -J[1][0] / det; K[1][1] = J[0][0] / det;
    } else {
        // cofactors (indices modulo 3)
        double C[3][3];
        #pragma unroll
        for (int i = 0; i < 3; ++i)
            #pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
                C[i][j] = J[i1][j1] * J[i2][j2] - J[i1][j2] * J[i2][j1];
            }
        det = J[0][0] * C[0][0] + J[0][1] * C[0][1] + J[0][2] * C[0][2];
        #pragma unroll
        for (int i = 0; i < 3; ++i)
            #pragma unroll
            for (int j = 0; j < 3; ++j) K[i][j] = C[j][i] / det;       // inverse = adj / det
    }
    // grad phi_a = K^T ghat_a, ghat_0 = -(1,..,1), ghat_{j+1} = e_j
    #pragma unroll
    for (int i = 0; i < D; ++i) {
        double s0 = 0.0;
        #pragma unroll
        for (int j = 0; j < D; ++j) { c.g[j + 1][i] = K[j][i]; s0 += K[j][i]; }
        c.g[0][i] = -s0;
    }
    c.detJ = fabs(det);
}

template <int D>
HEMO_HD void simplex_derive(SimplexCell<D>& c, const HemoForm& par) {
    constexpr int NV = D + 1;
    const double th = par.theta;
    #pragma unroll
    for (int a = 0; a < NV; ++a)
        #pragma unroll
        for (int k = 0; k < D; ++k) c.M[a][k] = th * c.U[a][k] + (1.0 - th) * c.N[a][k];
    #pragma unroll
    for (int i = 0; i < D; ++i) {
        #pragma unroll
        for (int j = 0; j < D; ++j) {
            double v = 0.0;
            #pragma unroll
            for (int a = 0; a < NV; ++a) v += c.g[a][i] * c.M[a][j];
            c.G[i][j] = v;
        }
        double v = 0.0;
        #pragma unroll
        for (int a = 0; a < NV; ++a) v += c.g[a][i] * c.P[a];
        c.gp[i] = v;
    }
    c.divu = 0.0;
    #pragma unroll
    for (int i = 0; i < D; ++i) c.divu += c.G[i][i];
    #pragma unroll
    for (int cc = 0; cc < NV; ++cc) {
        #pragma unroll
        for (int a = 0; a < NV; ++a) {
            double v = 0.0;
            #pragma unroll
            for (int i = 0; i < D; ++i) v += c.M[cc][i] * c.g[a][i];
            c.s[cc][a] = v;
        }
        #pragma unroll
        for (int k = 0; k < D; ++k) {
            double conv = 0.0;
            #pragma unroll
            for (int i = 0; i < D; ++i) conv += c.M[cc][i] * c.G[i][k];
            c.A[cc][k] = (par.a0 * c.U[cc][k] - c.H[cc][k]) * par.inv_dt + conv - c.fbody[k];
            c.R[cc][k] = par.rho * c.A[cc][k] + c.gp[k];
        }
    }
}

// Moments of one rule: T2[a][b] = |J| sum w tau phi_a phi_b, L0 = |J| sum w tau_lsic
template <int D>
HEMO_HD void simplex_moments(const SimplexCell<D>& c, const HemoForm& par, const SimplexRule<D>& r,
                             double T2[D + 1][D + 1], double& L0) {
    constexpr int NV = D + 1;
    const double h = c.h;
    const double inv_h2 = 1.0 / (h * h);
    const double t2inv = 2.0 * par.inv_dt, t3inv = 4.0 * par.nu * inv_h2;
    const double c23 = t2inv * t2inv + t3inv * t3inv;
    const double eps2 = par.eps0 * par.eps0;
    const double re_fac = h / (2.0 * par.nu);
    #pragma unroll
    for (int a = 0; a < NV; ++a)
        #pragma unroll
        for (int b = 0; b < NV; ++b) T2[a][b] = 0.0;
    L0 = 0.0;
    for (int q = 0; q < r.nq; ++q) {
        double v2 = 0.0;
        #pragma unroll
        for (int k = 0; k < D; ++k) {
            double u = 0.0;
            #pragma unroll
            for (int a = 0; a < NV; ++a) u += r.phi[q][a] * c.N[a][k];
            v2 += u * u;
        }
        const double t1 = fmax(4.0 * v2, eps2) * inv_h2;
        const double tau = 1.0 / sqrt(t1 + c23);
        const double v = sqrt(v2);
        const double Re = v * re_fac;
        const double z = (Re <= 3.0) ? Re / 3.0 : 1.0;
        const double wt = r.w[q] * tau;
        #pragma unroll
        for (int a = 0; a < NV; ++a)
            #pragma unroll
            for (int b = a; b < NV; ++b) T2[a][b] += wt * r.phi[q][a] * r.phi[q][b];
        L0 += r.w[q] * (0.5 * v * h * z);
    }
    #pragma unroll
    for (int a = 0; a < NV; ++a)
        #pragma unroll
        for (int b = a; b < NV; ++b) { T2[a][b] *= c.detJ; T2[b][a] = T2[a][b]; }
    L0 *= c.detJ;
}

// Element residual with the rules of the F_u (ru) and F_p (rp) block forms.
template <int D>
HEMO_HD void simplex_residual(const SimplexCell<D>& c, const HemoForm& par, const SimplexRule<D>& ru,
                              const SimplexRule<D>& rp, double Fu[D + 1][D], double Fp[D + 1]) {
    constexpr int NV = D + 1;
    const double rho = par.rho, mu = par.mu;
    double T2[NV][NV], L0, T2p[NV][NV], L0p, T1p[NV];
    simplex_moments<D>(c, par, ru, T2, L0);
    simplex_moments<D>(c, par, rp, T2p, L0p);
    #pragma unroll
    for (int d = 0; d < NV; ++d) {
        T1p[d] = 0.0;
        #pragma unroll
        for (int cc = 0; cc < NV; ++cc) T1p[d] += T2p[cc][d];
    }
    const double m0 = ru.m0 * c.detJ;
    double pbar = 0.0;
    #pragma unroll
    for (int b = 0; b < NV; ++b) pbar += ru.m1[b] * c.P[b];
    pbar *= c.detJ;
    #pragma unroll
    for (int a = 0; a < NV; ++a) {
        double Wd[NV];
        #pragma unroll
        for (int d = 0; d < NV; ++d) {
            double v = 0.0;
            #pragma unroll
            for (int cc = 0; cc < NV; ++cc) v += T2[cc][d] * c.s[cc][a];
            Wd[d] = v;
        }
        #pragma unroll
        for (int k = 0; k < D; ++k) {
            double v = 0.0;
            #pragma unroll
            for (int cc = 0; cc < NV; ++cc) v += rho * c.detJ * ru.m2[a][cc] * c.A[cc][k];
            double sg = 0.0;                                  // g_a . (2 mu eps)_{.k}
            #pragma unroll
            for (int i = 0; i < D; ++i) sg += c.g[a][i] * mu * (c.G[i][k] + c.G[k][i]);
            v += m0 * sg - c.g[a][k] * pbar;
            #pragma unroll
            for (int d = 0; d < NV; ++d) v += Wd[d] * c.R[d][k];
            v += L0 * rho * c.divu * c.g[a][k];
            Fu[a][k] = v;
        }
        double acc = 0.0;
        #pragma unroll
        for (int d = 0; d < NV; ++d) {
            double rg = 0.0;
            #pragma unroll
            for (int i = 0; i < D; ++i) rg += c.R[d][i] * c.g[a][i];
            acc += T1p[d] * rg;
        }
        Fp[a] = c.detJ * rp.m1[a] * c.divu + acc / rho;
    }
}

// Element Jacobian from the moments of the block rules: emit(a, b, ri, ci, value) with ri / ci in
// (u_0..u_{D-1}, p = D).  T2 / L0: moments of the J_uu rule; T1up / T1pu: column sums of T2 of the
// J_up / J_pu rules; T0pp: total of T2 of the J_pp rule; m0uu, m2uu, m1up, m1pu: polynomial moments.
template <int D, typename Emit>
HEMO_HD void simplex_jacobian_from_moments(const SimplexCell<D>& c, const HemoForm& par, const double T2[D + 1][D + 1],
                                           double L0, const double T1up[D + 1], const double T1pu[D + 1], double T0pp,
                                           double m0uu, const double m2uu[D + 1][D + 1], const double m1up[D + 1],
                                           const double m1pu[D + 1], Emit emit) {
    constexpr int NV = D + 1;
    const double rho = par.rho, mu = par.mu, idt = par.a0_dt, th = par.theta;
    const double m0 = m0uu * c.detJ;
    double RT[NV][D], Y[NV], V[NV];
    #pragma unroll
    for (int b = 0; b < NV; ++b) {
        #pragma unroll
        for (int k = 0; k < D; ++k) {
            double v = 0.0;
            #pragma unroll
            for (int d = 0; d < NV; ++d) v += T2[d][b] * c.R[d][k];
            RT[b][k] = v;
        }
        double y = 0.0, vv = 0.0;
        #pragma unroll
        for (int d = 0; d < NV; ++d) { y += T1pu[d] * c.s[d][b]; vv += T1up[d] * c.s[d][b]; }
        Y[b] = y; V[b] = vv;
    }
    #pragma unroll
    for (int a = 0; a < NV; ++a) {
        double TS[NV], Gga[D];
        #pragma unroll
        for (int cc = 0; cc < NV; ++cc) {
            double v = 0.0;
            #pragma unroll
            for (int d = 0; d < NV; ++d) v += T2[cc][d] * c.s[d][a];
            TS[cc] = v;
        }
        #pragma unroll
        for (int l = 0; l < D; ++l) {
            double v = 0.0;
            #pragma unroll
            for (int k = 0; k < D; ++k) v += c.G[l][k] * c.g[a][k];
            Gga[l] = v;
        }
        #pragma unroll
        for (int b = 0; b < NV; ++b) {
            const double m2ab = c.detJ * m2uu[a][b];
            double gab = 0.0, Zab = 0.0, Qab = 0.0;
            #pragma unroll
            for (int i = 0; i < D; ++i) gab += c.g[a][i] * c.g[b][i];
            const double Wab = TS[b];
            #pragma unroll
            for (int d = 0; d < NV; ++d) { Zab += TS[d] * c.s[d][b]; Qab += c.detJ * m2uu[a][d] * c.s[d][b]; }
            const double diag = rho * m2ab * idt + th * rho * Qab + th * mu * m0 * gab + rho * Wab * idt + th * rho * Zab;
            const double cG = th * rho * (m2ab + Wab);
            #pragma unroll
            for (int k = 0; k < D; ++k)
                #pragma unroll
                for (int l = 0; l < D; ++l) {
                    double v = cG * c.G[l][k] + th * mu * m0 * c.g[a][l] * c.g[b][k] + th * c.g[a][l] * RT[b][k] +
                               th * L0 * rho * c.g[a][k] * c.g[b][l];
                    if (k == l) v += diag;
                    emit(a, b, k, l, v);
                }
            const double m1b_up = c.detJ * m1up[b], m1a_pu = c.detJ * m1pu[a];
            #pragma unroll
            for (int k = 0; k < D; ++k) {
                emit(a, b, k, D, -m1b_up * c.g[a][k] + c.g[b][k] * V[a]);                                        // J_up
                emit(a, b, D, k, th * m1a_pu * c.g[b][k] + c.g[a][k] * (T1pu[b] * idt + th * Y[b]) + th * T1pu[b] * Gga[k]);   // J_pu
            }
            emit(a, b, D, D, T0pp / rho * gab);                                                                  // J_pp
        }
    }
}

// Column sums of a moment matrix
template <int D>
HEMO_HD void simplex_colsum(const double T[D + 1][D + 1], double T1[D + 1], double& T0) {
    T0 = 0.0;
    #pragma unroll
    for (int d = 0; d < D + 1; ++d) {
        T1[d] = 0.0;
        #pragma unroll
        for (int cc = 0; cc < D + 1; ++cc) T1[d] += T[cc][d];
        T0 += T1[d];
    }
}

// Element Jacobian with every block integrated by its own rule (ruu, rup, rpu, rpp).
template <int D, typename Emit>
HEMO_HD void simplex_jacobian(const SimplexCell<D>& c, const HemoForm& par, const SimplexRule<D>& ruu,
                              const SimplexRule<D>& rup, const SimplexRule<D>& rpu, const SimplexRule<D>& rpp,
                              Emit emit) {
    constexpr int NV = D + 1;
    double T2[NV][NV], L0, Tt[NV][NV], Lt, T1up[NV], T1pu[NV], T1[NV], T0, T0pp;
    simplex_moments<D>(c, par, rup, Tt, Lt);
    simplex_colsum<D>(Tt, T1up, T0);
    simplex_moments<D>(c, par, rpu, Tt, Lt);
    simplex_colsum<D>(Tt, T1pu, T0);
    simplex_moments<D>(c, par, rpp, Tt, Lt);
    simplex_colsum<D>(Tt, T1, T0pp);
    simplex_moments<D>(c, par, ruu, T2, L0);
    simplex_jacobian_from_moments<D>(c, par, T2, L0, T1up, T1pu, T0pp, ruu.m0, ruu.m2, rup.m1, rpu.m1, emit);
}


// ---------------------------------------------------------------------------
// Exterior-facet integrals on local facet lf (opposite vertex lf) of a simplex whose geometry and
// derived fields are set (simplex_geometry + simplex_derive).  Terms and coefficients: FacetSet in
// oracle/ns_oracle.py (stabilized_schur.py:79; stabilized_schur_pressure_backflow.py:192-217).
// The facet rule is given in barycentric coordinates of the facet's own vertices (ascending local
// order); weights sum to the reference facet measure (1 on an interval, 1/2 on a triangle).
//   res(a, k, value)          adds to F_u[a][k]
//   jac(a, b, k, ci, value)   adds to dF_u[a][k] / d(U[b][ci]) (ci < D) or / dP[b] (ci = D)
// ---------------------------------------------------------------------------
#define HEMO_SIMPLEX_MAXFQ 16

template <int D>
struct SimplexFacetRule {
    int nq;
    double lam[HEMO_SIMPLEX_MAXFQ][D];
    double w[HEMO_SIMPLEX_MAXFQ];
};

// pts: nq points on the reference facet ((D-1) coordinates each: s on [0,1]; (s,t) on the triangle)
template <int D>
inline void simplex_facet_rule_set(SimplexFacetRule<D>& r, const double* pts, const double* wts, int nq) {
    r.nq = nq;
    for (int q = 0; q < nq; ++q) {
        double s = 0.0;
        for (int j = 0; j + 1 < D; ++j) { r.lam[q][j + 1] = pts[(D - 1) * q + j]; s += pts[(D - 1) * q + j]; }
        r.lam[q][0] = 1.0 - s;
        r.w[q] = wts[q];
    }
}

// unit outward normal and physical / reference measure ratio of facet lf: the barycentric
// coordinate of the opposite vertex decreases towards the facet, so n = -grad phi_lf / |grad phi_lf|,
// and measure(F) = D |K| |grad phi_lf| with |K| = detJ / D!, reference measure 1 / (D-1)!.
template <int D>
HEMO_HD void simplex_facet_normal(const SimplexCell<D>& c, int lf, double nr[D], double& scale) {
    double g2 = 0.0;
    for (int i = 0; i < D; ++i) g2 += c.g[lf][i] * c.g[lf][i];
    const double gn = sqrt(g2);
    for (int i = 0; i < D; ++i) nr[i] = -c.g[lf][i] / gn;
    scale = c.detJ * gn;
}

template <int D, bool WANT_RES, bool WANT_JAC, typename Res, typename Jac>
HEMO_HD void simplex_facet(const SimplexCell<D>& c, const HemoForm& par, const hemo_facet_coef& co,
                           const SimplexFacetRule<D>& fr, int lf, Res res, Jac jac) {
    constexpr int NV = D + 1;
    const double th = par.theta, mu = par.mu, rho = par.rho;
    double nr[D], scale;
    simplex_facet_normal<D>(c, lf, nr, scale);
    double Pn[D][D];
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) Pn[i][j] = ((i == j) ? 1.0 : 0.0) - nr[i] * nr[j];
    // facet moments: Ph_a = int phi_a, Ph2_ab = int phi_a phi_b, B_ab = int (u_n.n)_- phi_a phi_b
    double Ph[NV], Ph2[NV][NV], B[NV][NV];
    for (int a = 0; a < NV; ++a) {
        Ph[a] = 0.0;
        for (int b = 0; b < NV; ++b) { Ph2[a][b] = 0.0; B[a][b] = 0.0; }
    }
    for (int q = 0; q < fr.nq; ++q) {
        double phi[NV];
        for (int j = 0, v = 0; v < NV; ++v) phi[v] = (v == lf) ? 0.0 : fr.lam[q][j++];
        const double w = fr.w[q] * scale;
        double unn = 0.0;
        for (int k = 0; k < D; ++k) {
            double u = 0.0;
            for (int a = 0; a < NV; ++a) u += phi[a] * c.N[a][k];
            unn += u * nr[k];
        }
        const double unm = 0.5 * (unn - fabs(unn));
        for (int a = 0; a < NV; ++a) {
            Ph[a] += w * phi[a];
            for (int b = 0; b < NV; ++b) {
                Ph2[a][b] += w * phi[a] * phi[b];
                B[a][b] += w * unm * phi[a] * phi[b];
            }
        }
    }
    double dn[NV], Png[NV][D];
    for (int a = 0; a < NV; ++a) {
        dn[a] = 0.0;
        for (int i = 0; i < D; ++i) dn[a] += c.g[a][i] * nr[i];
        for (int k = 0; k < D; ++k) {
            double v = 0.0;
            for (int i = 0; i < D; ++i) v += Pn[k][i] * c.g[a][i];
            Png[a][k] = v;
        }
    }
    const double pen = co.a_n * co.beta_n * mu / c.h;
    const double bf = co.a_b * co.beta_b * rho;
    if (WANT_JAC) {
        for (int a = 0; a < NV; ++a)
            for (int b = 0; b < NV; ++b)
                for (int k = 0; k < D; ++k) {
                    for (int l = 0; l < D; ++l) {
                        const double dkl = (k == l) ? 1.0 : 0.0;
                        double v = -th * co.a_g * mu * c.g[b][k] * nr[l] * Ph[a];
                        v -= th * co.a_s * mu * (c.g[b][k] * nr[l] + dn[b] * dkl) * Ph[a];
                        v -= th * co.a_n * mu * (Png[b][k] * nr[l] + dn[b] * Pn[k][l]) * Ph[a];
                        v -= th * co.a_n * mu * (Png[a][l] * nr[k] + dn[a] * Pn[k][l]) * Ph[b];
                        v += th * pen * Pn[k][l] * Ph2[a][b];
                        v -= th * bf * B[a][b] * dkl;
                        jac(a, b, k, l, v);
                    }
                    jac(a, b, k, D, co.a_p * nr[k] * Ph2[a][b]);
                }
    }
    if (WANT_RES) {
        double Gn[D], en[D], enT[D], Mi[D], MiT[D];
        for (int i = 0; i < D; ++i) {
            double gn = 0.0, e = 0.0;
            for (int j = 0; j < D; ++j) { gn += c.G[i][j] * nr[j]; e += 0.5 * (c.G[i][j] + c.G[j][i]) * nr[j]; }
            Gn[i] = gn; en[i] = e;
            double m = 0.0;
            for (int cc = 0; cc < NV; ++cc) m += Ph[cc] * c.M[cc][i];
            Mi[i] = m;
        }
        for (int k = 0; k < D; ++k) {
            double e = 0.0, m = 0.0;
            for (int i = 0; i < D; ++i) { e += Pn[k][i] * en[i]; m += Pn[k][i] * Mi[i]; }
            enT[k] = e; MiT[k] = m;
        }
        for (int a = 0; a < NV; ++a) {
            double pa = 0.0, Ma[D], Ba[D], gM = 0.0;
            for (int k = 0; k < D; ++k) { Ma[k] = 0.0; Ba[k] = 0.0; gM += c.g[a][k] * MiT[k]; }
            for (int b = 0; b < NV; ++b) {
                pa += Ph2[a][b] * c.P[b];
                for (int k = 0; k < D; ++k) { Ma[k] += Ph2[a][b] * c.M[b][k]; Ba[k] += B[a][b] * c.M[b][k]; }
            }
            for (int k = 0; k < D; ++k) {
                double MaT = 0.0;
                for (int i = 0; i < D; ++i) MaT += Pn[k][i] * Ma[i];
                double v = (co.a_p * pa + co.pconst * Ph[a]) * nr[k];
                v -= co.a_g * mu * Gn[k] * Ph[a];
                v -= co.a_s * 2.0 * mu * en[k] * Ph[a];
                v -= co.a_n * 2.0 * mu * enT[k] * Ph[a];
                v -= co.a_n * mu * (gM * nr[k] + dn[a] * MiT[k]);      // -(2 mu eps(v) n).u_T
                v += pen * MaT;
                v -= bf * Ba[k];
                res(a, k, v);
            }
        }
    }
}

// Q = int_F u_n . n ds on facet lf (exact for P1: measure times the mean of the facet's nodal values)
template <int D>
HEMO_HD double simplex_facet_flux(const SimplexCell<D>& c, int lf) {
    double nr[D], scale;
    simplex_facet_normal<D>(c, lf, nr, scale);
    double q = 0.0;
    for (int a = 0; a < D + 1; ++a) {
        if (a == lf) continue;
        for (int k = 0; k < D; ++k) q += c.N[a][k] * nr[k];
    }
    // measure = scale / (D-1)!, mean = sum / D
    return q * scale / (D == 3 ? 6.0 : 2.0);
}
