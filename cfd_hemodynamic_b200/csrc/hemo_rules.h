// Quadrature-rule tables shared by the device code and the host-compiled checks of the
// element routines (tests/host_q1): plain structs, no CUDA types.
#pragma once

#define HEMO_MAXFQ 8          // max facet quadrature points

// Constants of the weak form as the kernels see them: hemo_params (include/hemo.h) plus the time
// scheme of hemo_set_time_scheme.  The forms are evaluated at u_e = theta u + (1 - theta) u_n with
// the time derivative (a0 u - u_h) / dt; theta = 1/2, a0 = 1, u_h = u_n is the mid-point scheme of
// src/solvers/stabilized_schur.py:69-80, theta = 1 with a0 u - u_h = a0 u + a1 u_n + a2 u_nn the
// BDF scheme of src/solvers/stabilized_schur_bdf2.py:76-110.
struct HemoForm {
    double dt, rho, mu;
    double f[2];
    double eps0;
    double theta, a0;
    // derived on the host before every upload (hemo_form_finalize): divisions the point loops would repeat
    double inv_dt, a0_dt, inv_rho, nu;
};

static inline void hemo_form_finalize(HemoForm& f) {
    f.inv_dt = 1.0 / f.dt;
    f.a0_dt = f.a0 / f.dt;
    f.inv_rho = 1.0 / f.rho;
    f.nu = f.mu / f.rho;
}

struct HemoFacetRule {
    int nq;
    double s[HEMO_MAXFQ];
    double w[HEMO_MAXFQ];
};

// Cell rule of one block form on the reference quadrilateral [0,1]^2: points and weights as
// given (Basix: tensor Gauss-Jacobi, 12 x 12 points for the degree-22 forms); the Q1 basis
// and the bilinear geometry are evaluated from (xi, eta) on the fly.  The limit leaves room for
// 16 x 16 points (degree 30, where FFCx starts to warn): if UFL adds the degree of det J of the
// non-affine map to the estimate, the forms come out at degree 26 = 14 x 14 points.
#define HEMO_MAXQ_QUAD 256
struct HemoQuadRule {
    int nq;
    int alias;                      // lowest block id with an identical rule
    double pt[HEMO_MAXQ_QUAD][3];   // xi, eta, weight
};


// alias = lowest block id of the same group (residual forms F_u, F_p = ids 0, 1 | Jacobian forms
// J_uu..J_pp = ids 2..5) with an identical rule, so that a rule shared by several block forms is
// integrated once.
static inline void hemo_quad_rule_aliases(HemoQuadRule* rules, const bool* have, int nrules) {
    for (int b = 0; b < nrules; ++b) {
        if (!have[b]) continue;
        HemoQuadRule& rb = rules[b];
        rb.alias = b;
        const int first = (b <= 1) ? 0 : 2;
        for (int a = first; a < b; ++a) {
            if (!have[a] || rules[a].nq != rb.nq) continue;
            bool same = true;
            for (int q = 0; q < rb.nq && same; ++q)
                same = rules[a].pt[q][0] == rb.pt[q][0] && rules[a].pt[q][1] == rb.pt[q][1] &&
                       rules[a].pt[q][2] == rb.pt[q][2];
            if (same) { rb.alias = a; break; }
        }
    }
}
