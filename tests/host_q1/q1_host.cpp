// TEST INFRASTRUCTURE ONLY: compiles the __host__ __device__ element routines of
// cfd_hemodynamic_b200/csrc/q1_element.cuh with g++ so that their arithmetic can be checked
// against the numpy oracle without a GPU (tests/test_q1_host.py).  Not part of the product:
// libhemo_sm100.so never executes these routines on the CPU.
#include <stdint.h>
#include <string.h>

#include "../../cfd_hemodynamic_b200/csrc/q1_element.cuh"

static HemoForm g_par = {0, 0, 0, {0, 0}, 0, 0.5, 1.0, 0, 0, 0, 0};
static const double* g_uh = nullptr;   // history vector of the time derivative (null: u_n)

static void load(Q1Cell& cd, int c, const int32_t* cells, const double* x, const double* h, const double* sol,
                 const double* un, int n, int v[4]) {
    const double* uh = g_uh ? g_uh : un;
    for (int a = 0; a < 4; ++a) {
        v[a] = cells[4 * (int64_t)c + a];
        cd.X[a][0] = x[2 * v[a]]; cd.X[a][1] = x[2 * v[a] + 1];
        cd.U[a][0] = sol[2 * v[a]]; cd.U[a][1] = sol[2 * v[a] + 1];
        cd.N[a][0] = un[2 * v[a]]; cd.N[a][1] = un[2 * v[a] + 1];
        cd.H[a][0] = uh[2 * v[a]]; cd.H[a][1] = uh[2 * v[a] + 1];
        cd.P[a] = sol[2 * (int64_t)n + v[a]];
    }
    cd.h = h[c];
    hemo_form_finalize(g_par);
    q1_prepare(cd, g_par);
}

static HemoQuadRule g_rules[6];
static HemoFacetRule g_frule;

extern "C" {

void q1h_set_rule(int block, const double* pts, const double* wts, int nq) {
    static bool have[6] = {false, false, false, false, false, false};
    HemoQuadRule& r = g_rules[block];
    r.nq = nq;
    for (int q = 0; q < nq; ++q) { r.pt[q][0] = pts[2 * q]; r.pt[q][1] = pts[2 * q + 1]; r.pt[q][2] = wts[q]; }
    have[block] = true;
    hemo_quad_rule_aliases(g_rules, have, 6);
}

void q1h_set_facet_rule(const double* s, const double* w, int nq) {
    g_frule.nq = nq;
    for (int q = 0; q < nq; ++q) { g_frule.s[q] = s[q]; g_frule.w[q] = w[q]; }
}

void q1h_set_params(const hemo_params* p) {
    g_par.dt = p->dt; g_par.rho = p->rho; g_par.mu = p->mu;
    g_par.f[0] = p->f[0]; g_par.f[1] = p->f[1]; g_par.eps0 = p->eps0;
}

void q1h_set_time_scheme(double theta, double a0, const double* uh) { g_par.theta = theta; g_par.a0 = a0; g_uh = uh; }

int q1h_alias(int block) { return g_rules[block].alias; }

// Ae: SoA [(a*4+b)*9 + ri*3+ci][E] like the device buffer
void q1h_cell_jacobian(int E, int n, const int32_t* cells, const double* x, const double* h, const double* sol,
                       const double* un, double* Ae) {
    for (int c = 0; c < E; ++c) {
        Q1Cell cd; int v[4];
        load(cd, c, cells, x, h, sol, un, n, v);
        auto emit = [&](int slot, double val) { Ae[(int64_t)slot * E + c] = val; };
        q1_cell_jacobian_uu<0>(cd, g_par, g_rules, emit);
        q1_cell_jacobian_uu<2>(cd, g_par, g_rules, emit);
        q1_cell_jacobian_p(cd, g_par, g_rules, emit);
    }
}

// Fe: SoA [a*3+comp][E]; dvec: 3n lifting vector or NULL
void q1h_cell_residual(int E, int n, const int32_t* cells, const double* x, const double* h, const double* sol,
                       const double* un, const uint8_t* cellflag, const double* dvec, double* Fe) {
    for (int c = 0; c < E; ++c) {
        Q1Cell cd; int v[4];
        load(cd, c, cells, x, h, sol, un, n, v);
        double Fu[4][2], Fp[4];
        q1_cell_residual(cd, g_par, g_rules, Fu, Fp);
        if (cellflag && cellflag[c]) {
            double dl[4][3];
            for (int b = 0; b < 4; ++b) {
                dl[b][0] = dvec[2 * v[b]]; dl[b][1] = dvec[2 * v[b] + 1]; dl[b][2] = dvec[2 * (int64_t)n + v[b]];
            }
            q1_cell_lift(cd, g_par, g_rules, dl, Fu, Fp);
        }
        for (int a = 0; a < 4; ++a) {
            Fe[(int64_t)(a * 3 + 0) * E + c] = Fu[a][0];
            Fe[(int64_t)(a * 3 + 1) * E + c] = Fu[a][1];
            Fe[(int64_t)(a * 3 + 2) * E + c] = Fp[a];
        }
    }
}

// mode 1: += into Ae; mode 0: += into Fe (with lifting when cellflag is given)
void q1h_facets(int mode, int m, int E, int n, const int32_t* fcells, const int32_t* fmask, const hemo_facet_coef* co,
                const int32_t* cells, const double* x, const double* h, const double* sol, const double* un,
                const uint8_t* cellflag, const double* dvec, double* out) {
    for (int t = 0; t < m; ++t) {
        const int c = fcells[t];
        Q1Cell cd; int v[4];
        load(cd, c, cells, x, h, sol, un, n, v);
        if (mode == 1) {
            q1_cell_facets(cd, g_par, g_frule, *co, fmask[t], false, true, [&](int, int, double) {},
                           [&](int a, int b, int ri, int ci, double val) {
                               out[((int64_t)(a * 4 + b) * 9 + ri * 3 + ci) * E + c] += val;
                           });
        } else {
            double dl[4][3];
            bool lift = false;
            if (cellflag && cellflag[c]) {
                for (int b = 0; b < 4; ++b) {
                    dl[b][0] = dvec[2 * v[b]]; dl[b][1] = dvec[2 * v[b] + 1]; dl[b][2] = dvec[2 * (int64_t)n + v[b]];
                    lift = lift || dl[b][0] != 0.0 || dl[b][1] != 0.0 || dl[b][2] != 0.0;
                }
            }
            q1_cell_facets(cd, g_par, g_frule, *co, fmask[t], true, lift,
                           [&](int a, int k, double val) { out[(int64_t)(a * 3 + k) * E + c] += val; },
                           [&](int a, int b, int ri, int ci, double val) {
                               out[(int64_t)(a * 3 + ri) * E + c] += val * dl[b][ci];
                           });
        }
    }
}

double q1h_flux(int m, const int32_t* fcells, const int32_t* fmask, const int32_t* cells, const double* x,
                const double* un) {
    double q = 0.0;
    for (int t = 0; t < m; ++t) {
        Q1Cell cd;
        for (int a = 0; a < 4; ++a) {
            const int v = cells[4 * (int64_t)fcells[t] + a];
            cd.X[a][0] = x[2 * v]; cd.X[a][1] = x[2 * v + 1];
            cd.N[a][0] = un[2 * v]; cd.N[a][1] = un[2 * v + 1];
        }
        q += q1_cell_flux(cd, fmask[t]);
    }
    return q;
}

void q1h_laplace_mass(int E, const int32_t* cells, const double* x, double* Ke, double* Me) {
    for (int c = 0; c < E; ++c) {
        Q1Cell cd;
        for (int a = 0; a < 4; ++a) {
            const int v = cells[4 * (int64_t)c + a];
            cd.X[a][0] = x[2 * v]; cd.X[a][1] = x[2 * v + 1];
        }
        double K[4][4], M[4];
        q1_cell_laplace_mass(cd, K, M);
        for (int a = 0; a < 4; ++a) {
            for (int b = 0; b < 4; ++b) Ke[(int64_t)(a * 4 + b) * E + c] = K[a][b];
            Me[(int64_t)a * E + c] = M[a];
        }
    }
}

}  // extern "C"
