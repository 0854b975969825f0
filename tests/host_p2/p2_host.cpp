// TEST INFRASTRUCTURE ONLY: compiles the host+device element routines of cfd_hemodynamic_b200/csrc/p2_element.cuh with
// g++ so that their arithmetic can be checked against oracle/pk_oracle.py without a GPU (tests/test_p2_host.py), through
// the same SoA element buffers the CUDA kernels of assembly_p2.cu write.  libhemo_sm100.so never runs these on the CPU.
#include <stdint.h>
#include <string.h>

#include "../../cfd_hemodynamic_b200/csrc/p2_element.cuh"

static HemoForm g_par = {0, 0, 0, {0, 0}, 0, 0.5, 1.0, 0, 0, 0, 0};
static HemoP2Rule g_rules[6];
static bool g_have[6] = {false, false, false, false, false, false};
static HemoFacetRule g_frule;
static const double* g_uh = nullptr;

static void load(P2Cell& cd, int c, const int32_t* cells, const double* x, const double* h, const double* sol, const double* un,
                 int n, int v[6]) {
    const double* uh = g_uh ? g_uh : un;
    for (int a = 0; a < 6; ++a) {
        v[a] = cells[6 * (int64_t)c + a];
        if (a < 3) { cd.X[a][0] = x[2 * v[a]]; cd.X[a][1] = x[2 * v[a] + 1]; }
        cd.U[a][0] = sol[2 * v[a]]; cd.U[a][1] = sol[2 * v[a] + 1];
        cd.N[a][0] = un[2 * v[a]]; cd.N[a][1] = un[2 * v[a] + 1];
        cd.H[a][0] = uh[2 * v[a]]; cd.H[a][1] = uh[2 * v[a] + 1];
        cd.P[a] = sol[2 * (int64_t)n + v[a]];
    }
    cd.h = h[c];
    hemo_form_finalize(g_par);
    p2_prepare(cd, g_par);
}

extern "C" {

void p2h_set_rule(int block, const double* pts, const double* wts, int nq) {
    HemoP2Rule& r = g_rules[block];
    r.nq = nq;
    for (int q = 0; q < nq; ++q) { r.pt[q][0] = pts[2 * q]; r.pt[q][1] = pts[2 * q + 1]; r.pt[q][2] = wts[q]; }
    g_have[block] = true;
    hemo_p2_rule_aliases(g_rules, g_have, 6);
}
void p2h_set_facet_rule(const double* s, const double* w, int nq) {
    g_frule.nq = nq;
    for (int q = 0; q < nq; ++q) { g_frule.s[q] = s[q]; g_frule.w[q] = w[q]; }
}
void p2h_set_params(const hemo_params* p) {
    g_par.dt = p->dt; g_par.rho = p->rho; g_par.mu = p->mu;
    g_par.f[0] = p->f[0]; g_par.f[1] = p->f[1]; g_par.eps0 = p->eps0;
}
void p2h_set_time_scheme(double theta, double a0, const double* uh) { g_par.theta = theta; g_par.a0 = a0; g_uh = uh; }
int p2h_alias(int block) { return g_rules[block].alias; }

// Ae: SoA [(a*6+b)*9 + ri*3+ci][E]
void p2h_cell_jacobian(int E, int n, const int32_t* cells, const double* x, const double* h, const double* sol, const double* un,
                       double* Ae) {
    for (int c = 0; c < E; ++c) {
        P2Cell cd;
        int v[6];
        load(cd, c, cells, x, h, sol, un, n, v);
        for (int a = 0; a < 6; ++a)
            p2_cell_jacobian_rows(cd, g_par, g_rules, a, [&](int slot, double val) { Ae[(int64_t)slot * E + c] = val; });
    }
}

// Fe: SoA [a*3+comp][E]
void p2h_cell_residual(int E, int n, const int32_t* cells, const double* x, const double* h, const double* sol, const double* un,
                       double* Fe) {
    for (int c = 0; c < E; ++c) {
        P2Cell cd;
        int v[6];
        load(cd, c, cells, x, h, sol, un, n, v);
        double Fu[6][2], Fp[6];
        p2_cell_residual(cd, g_par, g_rules, Fu, Fp);
        for (int a = 0; a < 6; ++a) {
            Fe[(int64_t)(a * 3 + 0) * E + c] = Fu[a][0];
            Fe[(int64_t)(a * 3 + 1) * E + c] = Fu[a][1];
            Fe[(int64_t)(a * 3 + 2) * E + c] = Fp[a];
        }
    }
}

// adds the facet terms of the m boundary cells into Fe (mode 0) or Ae (mode 1)
void p2h_facets(int mode, int m, const int32_t* fcells, const int32_t* fmask, const hemo_facet_coef* co, int E, int n,
                const int32_t* cells, const double* x, const double* h, const double* sol, const double* un, double* out) {
    for (int t = 0; t < m; ++t) {
        const int c = fcells[t];
        P2Cell cd;
        int v[6];
        load(cd, c, cells, x, h, sol, un, n, v);
        if (mode == 0) {
            double Fu[6][2];
            p2_facet_residual(cd, g_par, g_frule, *co, fmask[t], Fu);
            for (int a = 0; a < 6; ++a) {
                out[(int64_t)(a * 3 + 0) * E + c] += Fu[a][0];
                out[(int64_t)(a * 3 + 1) * E + c] += Fu[a][1];
            }
        } else {
            for (int b = 0; b < 6; ++b)
                for (int ci = 0; ci < 3; ++ci) {
                    double col[6][2];
                    p2_facet_column(cd, g_par, g_frule, *co, fmask[t], b, ci, col);
                    for (int a = 0; a < 6; ++a)
                        for (int ri = 0; ri < 2; ++ri) out[(int64_t)((a * 6 + b) * 9 + ri * 3 + ci) * E + c] += col[a][ri];
                }
        }
    }
}

double p2h_flux(int m, const int32_t* fcells, const int32_t* fmask, const int32_t* cells, const double* x, const double* un) {
    double total = 0.0;
    for (int t = 0; t < m; ++t) {
        P2Cell cd;
        const int c = fcells[t];
        for (int a = 0; a < 6; ++a) {
            const int v = cells[6 * (int64_t)c + a];
            if (a < 3) { cd.X[a][0] = x[2 * v]; cd.X[a][1] = x[2 * v + 1]; }
            cd.N[a][0] = un[2 * v]; cd.N[a][1] = un[2 * v + 1];
        }
        total += p2_cell_flux(cd, g_frule, fmask[t]);
    }
    return total;
}

void p2h_laplace_mass(int E, const int32_t* cells, const double* x, double* Ke /*[36][E]*/, double* Me /*[6][E]*/) {
    for (int c = 0; c < E; ++c) {
        P2Cell cd;
        memset(&cd, 0, sizeof cd);
        for (int a = 0; a < 3; ++a) {
            const int v = cells[6 * (int64_t)c + a];
            cd.X[a][0] = x[2 * v]; cd.X[a][1] = x[2 * v + 1];
        }
        cd.h = 1.0;
        HemoForm par = g_par;
        par.dt = par.rho = par.mu = 1.0;
        hemo_form_finalize(par);
        p2_prepare(cd, par);
        double K[6][6], M[6];
        p2_cell_laplace_mass(cd, K, M);
        for (int a = 0; a < 6; ++a) {
            for (int b = 0; b < 6; ++b) Ke[(int64_t)(a * 6 + b) * E + c] = K[a][b];
            Me[(int64_t)a * E + c] = M[a];
        }
    }
}
}
