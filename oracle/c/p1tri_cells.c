/* CPU ORACLE (C restatement) — TEST / BASELINE INFRASTRUCTURE ONLY, never linked into the product.
 *
 * FFCx-style cell kernels for the stabilized Navier–Stokes forms on P1–P1 triangles: the full
 * integrand is evaluated at every quadrature point of each block form's rule (no moment
 * factorisation), one OpenMP thread team over cells.  Follows
 *   /root/reference/src/solvers/stabilized_schur.py:60-123  (residual F)
 *   /root/reference/src/solvers/stabilized_schur.py:185-189 (J = derivative(F), per-block forms)
 * and is checked against oracle/ns_oracle.py (tests/test_oracle.py).  PARITY UNPINNED against the real
 * DOLFINx/FFCx kernels (not installable here).
 *
 * Local dof order: u(a,k) -> 2a+k, p(a) -> 6+a.  Ae: E x 9 x 9 row-major, Fe: E x 9.
 * rule r: 0 Fu, 1 Fp, 2 uu, 3 up, 4 pu, 5 pp; pts[r] = nq[r] x 2, wts[r] = nq[r].
 */
#include <math.h>
#include <string.h>

typedef struct {
    double g[3][2], det, U[3][2], N[3][2], P[3], h;
    double M[3][2], G[2][2], gp[2], divu;
} cell_t;

static void load_cell(cell_t* c, int e, const double* x, const int* cells, const double* h, const double* u,
                      const double* p, const double* un) {
    double X[3][2];
    for (int a = 0; a < 3; ++a) {
        const int v = cells[3 * e + a];
        X[a][0] = x[2 * v]; X[a][1] = x[2 * v + 1];
        c->U[a][0] = u[2 * v]; c->U[a][1] = u[2 * v + 1];
        c->N[a][0] = un[2 * v]; c->N[a][1] = un[2 * v + 1];
        c->P[a] = p[v];
        c->M[a][0] = 0.5 * (c->U[a][0] + c->N[a][0]);
        c->M[a][1] = 0.5 * (c->U[a][1] + c->N[a][1]);
    }
    const double j00 = X[1][0] - X[0][0], j01 = X[2][0] - X[0][0];
    const double j10 = X[1][1] - X[0][1], j11 = X[2][1] - X[0][1];
    const double det = j00 * j11 - j01 * j10;
    const double i00 = j11 / det, i01 = -j01 / det, i10 = -j10 / det, i11 = j00 / det;
    c->g[1][0] = i00; c->g[1][1] = i01; c->g[2][0] = i10; c->g[2][1] = i11;
    c->g[0][0] = -(i00 + i10); c->g[0][1] = -(i01 + i11);
    c->det = fabs(det);
    c->h = h[e];
    for (int i = 0; i < 2; ++i) {
        for (int j = 0; j < 2; ++j)
            c->G[i][j] = c->g[0][i] * c->M[0][j] + c->g[1][i] * c->M[1][j] + c->g[2][i] * c->M[2][j];
        c->gp[i] = c->g[0][i] * c->P[0] + c->g[1][i] * c->P[1] + c->g[2][i] * c->P[2];
    }
    c->divu = c->G[0][0] + c->G[1][1];
}

typedef struct {
    double phi[3], w, u[2], un[2], um[2], tau, tau_l, conv[2], R[2], umd[3], dudt[2], p;
} point_t;

static void eval_point(point_t* q, const cell_t* c, double xi, double eta, double wt, double dt, double rho, double mu,
                       const double* f, double eps0) {
    q->phi[0] = 1.0 - xi - eta; q->phi[1] = xi; q->phi[2] = eta;
    q->w = wt * c->det;
    q->p = 0.0;
    for (int k = 0; k < 2; ++k) {
        q->u[k] = q->phi[0] * c->U[0][k] + q->phi[1] * c->U[1][k] + q->phi[2] * c->U[2][k];
        q->un[k] = q->phi[0] * c->N[0][k] + q->phi[1] * c->N[1][k] + q->phi[2] * c->N[2][k];
        q->um[k] = 0.5 * (q->u[k] + q->un[k]);
        q->dudt[k] = (q->u[k] - q->un[k]) / dt;
    }
    for (int a = 0; a < 3; ++a) q->p += q->phi[a] * c->P[a];
    const double nu = mu / rho;
    const double vnorm = sqrt(q->un[0] * q->un[0] + q->un[1] * q->un[1]);
    const double two_v = 2.0 * vnorm;
    const double t1 = c->h / (two_v >= eps0 ? two_v : eps0);
    const double t2 = dt / 2.0;
    const double t3 = (c->h * c->h) / (4.0 * nu);
    q->tau = 1.0 / sqrt(1.0 / (t1 * t1) + 1.0 / (t2 * t2) + 1.0 / (t3 * t3));   /* (...)**(-0.5), :100-108 */
    const double Re = (vnorm * c->h) / (2.0 * nu);
    const double z = (Re <= 3.0) ? Re / 3.0 : 1.0;
    q->tau_l = (vnorm * c->h * z) / 2.0;
    for (int k = 0; k < 2; ++k) {
        q->conv[k] = q->um[0] * c->G[0][k] + q->um[1] * c->G[1][k];
        q->R[k] = rho * (q->dudt[k] + q->conv[k]) + c->gp[k] - rho * f[k];
    }
    for (int a = 0; a < 3; ++a) q->umd[a] = q->um[0] * c->g[a][0] + q->um[1] * c->g[a][1];
}

void hemo_ref_cells(int E, const double* x, const int* cells, const double* h, const double* u, const double* p,
                    const double* un, double dt, double rho, double mu, const double* f, double eps0, const int* nq,
                    const double* const* pts, const double* const* wts, double* Ae, double* Fe) {
#pragma omp parallel for schedule(static)
    for (int e = 0; e < E; ++e) {
        cell_t c;
        point_t q;
        load_cell(&c, e, x, cells, h, u, p, un);
        if (Fe) {
            double* F = Fe + 9 * (long)e;
            memset(F, 0, 9 * sizeof(double));
            const double eps00 = c.G[0][0], eps11 = c.G[1][1], eps01 = 0.5 * (c.G[0][1] + c.G[1][0]);
            for (int iq = 0; iq < nq[0]; ++iq) {           /* F_u, rule 0 */
                eval_point(&q, &c, pts[0][2 * iq], pts[0][2 * iq + 1], wts[0][iq], dt, rho, mu, f, eps0);
                const double s[2][2] = {{2 * mu * eps00 - q.p, 2 * mu * eps01}, {2 * mu * eps01, 2 * mu * eps11 - q.p}};
                for (int a = 0; a < 3; ++a)
                    for (int k = 0; k < 2; ++k)
                        F[2 * a + k] += q.w * (rho * q.phi[a] * (q.dudt[k] + q.conv[k] - f[k]) +
                                               c.g[a][0] * s[0][k] + c.g[a][1] * s[1][k] + q.tau * q.umd[a] * q.R[k] +
                                               q.tau_l * rho * c.divu * c.g[a][k]);
            }
            for (int iq = 0; iq < nq[1]; ++iq) {           /* F_p, rule 1 */
                eval_point(&q, &c, pts[1][2 * iq], pts[1][2 * iq + 1], wts[1][iq], dt, rho, mu, f, eps0);
                for (int a = 0; a < 3; ++a)
                    F[6 + a] += q.w * (q.phi[a] * c.divu + (q.tau / rho) * (q.R[0] * c.g[a][0] + q.R[1] * c.g[a][1]));
            }
        }
        if (Ae) {
            double* A = Ae + 81 * (long)e;
            memset(A, 0, 81 * sizeof(double));
            double dd[3][3];
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) dd[a][b] = c.g[a][0] * c.g[b][0] + c.g[a][1] * c.g[b][1];
            /* one loop nest per block form (each has its own rule, stabilized_schur.py:188-189); point-only
             * factors are hoisted out of the (a, b) loops the way FFCx hoists them */
            for (int iq = 0; iq < nq[2]; ++iq) {               /* J_uu */
                eval_point(&q, &c, pts[2][2 * iq], pts[2][2 * iq + 1], wts[2][iq], dt, rho, mu, f, eps0);
                for (int b = 0; b < 3; ++b) {
                    const double s = rho * (q.phi[b] / dt + 0.5 * q.umd[b]);
                    double dR[2][2];                           /* dR[l][k] = d R_k / d u_(b,l) */
                    for (int l = 0; l < 2; ++l)
                        for (int k = 0; k < 2; ++k) dR[l][k] = s * (k == l) + 0.5 * rho * q.phi[b] * c.G[l][k];
                    for (int a = 0; a < 3; ++a) {
                        const double ta = q.phi[a] + q.tau * q.umd[a];
                        for (int k = 0; k < 2; ++k)
                            for (int l = 0; l < 2; ++l) {
                                const double visc = 0.5 * mu * (dd[a][b] * (k == l) + c.g[a][l] * c.g[b][k]);
                                A[(2 * a + k) * 9 + 2 * b + l] += q.w * (ta * dR[l][k] + visc +
                                    0.5 * q.tau * q.R[k] * q.phi[b] * c.g[a][l] + 0.5 * q.tau_l * rho * c.g[a][k] * c.g[b][l]);
                            }
                    }
                }
            }
            for (int iq = 0; iq < nq[3]; ++iq) {               /* J_up */
                eval_point(&q, &c, pts[3][2 * iq], pts[3][2 * iq + 1], wts[3][iq], dt, rho, mu, f, eps0);
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b)
                        for (int k = 0; k < 2; ++k)
                            A[(2 * a + k) * 9 + 6 + b] += q.w * (-q.phi[b] * c.g[a][k] + q.tau * q.umd[a] * c.g[b][k]);
            }
            for (int iq = 0; iq < nq[4]; ++iq) {               /* J_pu */
                eval_point(&q, &c, pts[4][2 * iq], pts[4][2 * iq + 1], wts[4][iq], dt, rho, mu, f, eps0);
                for (int b = 0; b < 3; ++b) {
                    const double s = rho * (q.phi[b] / dt + 0.5 * q.umd[b]);
                    for (int l = 0; l < 2; ++l) {
                        const double dR0 = s * (l == 0) + 0.5 * rho * q.phi[b] * c.G[l][0];
                        const double dR1 = s * (l == 1) + 0.5 * rho * q.phi[b] * c.G[l][1];
                        for (int a = 0; a < 3; ++a)
                            A[(6 + a) * 9 + 2 * b + l] += q.w * (0.5 * q.phi[a] * c.g[b][l] +
                                (q.tau / rho) * (dR0 * c.g[a][0] + dR1 * c.g[a][1]));
                    }
                }
            }
            for (int iq = 0; iq < nq[5]; ++iq) {               /* J_pp */
                eval_point(&q, &c, pts[5][2 * iq], pts[5][2 * iq + 1], wts[5][iq], dt, rho, mu, f, eps0);
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) A[(6 + a) * 9 + 6 + b] += q.w * q.tau / rho * dd[a][b];
            }
        }
    }
}
