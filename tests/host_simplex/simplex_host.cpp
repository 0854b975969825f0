// TEST INFRASTRUCTURE ONLY: compiles cfd_hemodynamic_b200/csrc/simplex_element.cuh with g++ so the
// dimension-generic P1 simplex routines (triangles, tetrahedra) can be checked against
// oracle/simplex_oracle.py without a GPU (tests/test_simplex_host.py).
#include <stdint.h>

#include "../../cfd_hemodynamic_b200/csrc/simplex_element.cuh"

static HemoForm g_par = {0, 0, 0, {0, 0}, 0, 0.5, 1.0, 0, 0, 0, 0};
static double g_f[3] = {0, 0, 0};

template <int D>
struct Rules {
    SimplexRule<D> r[6];
};
static Rules<2> g_r2;
static Rules<3> g_r3;

template <int D>
static void run(Rules<D>& R, int E, int n, const int32_t* cells, const double* x, const double* h, const double* sol,
                const double* un, const double* uh, double* Ae, double* Fe) {
    constexpr int NV = D + 1, B = D + 1;        // B: scalars per node (u components + p)
    hemo_form_finalize(g_par);
    for (int c = 0; c < E; ++c) {
        SimplexCell<D> cd;
        double X[NV][D];
        for (int a = 0; a < NV; ++a) {
            const int v = cells[NV * (int64_t)c + a];
            for (int k = 0; k < D; ++k) {
                X[a][k] = x[D * (int64_t)v + k];
                cd.U[a][k] = sol[D * (int64_t)v + k];
                cd.N[a][k] = un[D * (int64_t)v + k];
                cd.H[a][k] = uh[D * (int64_t)v + k];
            }
            cd.P[a] = sol[D * (int64_t)n + v];
        }
        for (int k = 0; k < D; ++k) cd.fbody[k] = g_f[k];
        cd.h = h[c];
        simplex_geometry<D>(cd, X);
        simplex_derive<D>(cd, g_par);
        double Fu[NV][D], Fp[NV];
        simplex_residual<D>(cd, g_par, R.r[0], R.r[1], Fu, Fp);
        for (int a = 0; a < NV; ++a) {
            for (int k = 0; k < D; ++k) Fe[((int64_t)c * NV + a) * B + k] = Fu[a][k];
            Fe[((int64_t)c * NV + a) * B + D] = Fp[a];
        }
        simplex_jacobian<D>(cd, g_par, R.r[2], R.r[3], R.r[4], R.r[5], [&](int a, int b, int ri, int ci, double v) {
            Ae[((((int64_t)c * NV + a) * NV + b) * B + ri) * B + ci] = v;
        });
    }
}

extern "C" {

void sxh_set_params(double dt, double rho, double mu, const double* f, double eps0, double theta, double a0) {
    g_par.dt = dt; g_par.rho = rho; g_par.mu = mu; g_par.eps0 = eps0; g_par.theta = theta; g_par.a0 = a0;
    g_par.f[0] = f[0]; g_par.f[1] = f[1];
    g_f[0] = f[0]; g_f[1] = f[1]; g_f[2] = f[2];
}

int sxh_set_rule(int dim, int block, const double* pts, const double* wts, int nq) {
    if (nq > HEMO_SIMPLEX_MAXQ) return -1;
    if (dim == 2) simplex_rule_set<2>(g_r2.r[block], pts, wts, nq);
    else simplex_rule_set<3>(g_r3.r[block], pts, wts, nq);
    return 0;
}

// Ae: [E][NV][NV][B][B], Fe: [E][NV][B] with B = dim + 1 (u components, then p)
void sxh_cells(int dim, int E, int n, const int32_t* cells, const double* x, const double* h, const double* sol,
               const double* un, const double* uh, double* Ae, double* Fe) {
    if (dim == 2) run<2>(g_r2, E, n, cells, x, h, sol, un, uh, Ae, Fe);
    else run<3>(g_r3, E, n, cells, x, h, sol, un, uh, Ae, Fe);
}

}  // extern "C"
