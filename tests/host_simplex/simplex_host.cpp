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

// ---------------------------------------------------------------------------
// Exterior-facet routine (any D) and the emulated tetrahedron assembly pipeline: the per-thread
// items of csrc/tet_items.cuh walked over their index ranges in the order the kernels of
// assembly_tet.cu are launched.
// ---------------------------------------------------------------------------
#include "../../cfd_hemodynamic_b200/csrc/tet_items.cuh"

template <int D>
static void run_facets(int m, int n, const int32_t* pairs, const int32_t* cells, const double* x, const double* h,
                       const double* sol, const double* un, const hemo_facet_coef* co, const double* fpts,
                       const double* fwts, int nq, double* Fu, double* J, double* flux) {
    constexpr int NV = D + 1, B = D + 1;
    hemo_form_finalize(g_par);
    SimplexFacetRule<D> fr;
    simplex_facet_rule_set<D>(fr, fpts, fwts, nq);
    for (int t = 0; t < m; ++t) {
        const int c = pairs[2 * t], lf = pairs[2 * t + 1];
        SimplexCell<D> cd;
        double X[NV][D];
        for (int a = 0; a < NV; ++a) {
            const int v = cells[NV * (int64_t)c + a];
            for (int k = 0; k < D; ++k) {
                X[a][k] = x[D * (int64_t)v + k];
                cd.U[a][k] = sol[D * (int64_t)v + k];
                cd.N[a][k] = un[D * (int64_t)v + k];
                cd.H[a][k] = cd.N[a][k];
            }
            cd.P[a] = sol[D * (int64_t)n + v];
        }
        for (int k = 0; k < D; ++k) cd.fbody[k] = 0.0;
        cd.h = h[c];
        simplex_geometry<D>(cd, X);
        simplex_derive<D>(cd, g_par);
        double* fu = Fu + (int64_t)t * NV * D;
        double* jj = J + (int64_t)t * NV * NV * D * B;
        simplex_facet<D, true, true>(
            cd, g_par, *co, fr, lf, [&](int a, int k, double v) { fu[a * D + k] += v; },
            [&](int a, int b, int k, int ci, double v) { jj[((a * NV + b) * D + k) * B + ci] += v; });
        flux[t] = simplex_facet_flux<D>(cd, lf);
    }
}

extern "C" {

// Fu: [m][NV][D], J: [m][NV][NV][D][D+1] (zero-initialised by the caller), flux: [m]
void sxh_facets(int dim, int m, int n, const int32_t* pairs, const int32_t* cells, const double* x, const double* h,
                const double* sol, const double* un, const double* coef8, const double* fpts, const double* fwts, int nq,
                double* Fu, double* J, double* flux) {
    hemo_facet_coef co = {coef8[0], coef8[1], coef8[2], coef8[3], coef8[4], coef8[5], coef8[6], coef8[7]};
    if (dim == 2) run_facets<2>(m, n, pairs, cells, x, h, sol, un, &co, fpts, fwts, nq, Fu, J, flux);
    else run_facets<3>(m, n, pairs, cells, x, h, sol, un, &co, fpts, fwts, nq, Fu, J, flux);
}

// Emulated launch sequence of hemo_tet_assemble_jacobian / hemo_tet_assemble_residual / hemo_tet_spmv /
// hemo_outlet_flux on tetrahedra.  Element buffers SoA as on the device.  dofflag may be null (no
// Dirichlet conditions).  Outputs: vals (16 nnz_node), b (4n), y = J xv (4n), flux (1).
void txh_assemble(int E, int n, int64_t nnz_node, const int32_t* cells, const double* x, const double* h,
                  const double* sol, const double* un, const double* uh, const int32_t* nrowptr, const int32_t* ncol,
                  const int32_t* rowof, const int32_t* mseg_ptr, const int32_t* mseg_src, const int32_t* vseg_ptr,
                  const int32_t* vseg_src, int m, const int32_t* fcells, const int32_t* fmask, const double* coef8,
                  const double* fpts, const double* fwts, int nfq, const uint8_t* dofflag, const double* dofmult,
                  const uint8_t* cellflag, const double* g, const double* xv, double* Ae, double* Fe, double* vals,
                  double* b, double* y, double* flux) {
    hemo_form_finalize(g_par);
    hemo_facet_coef co = {coef8[0], coef8[1], coef8[2], coef8[3], coef8[4], coef8[5], coef8[6], coef8[7]};
    SimplexFacetRule<3> fr;
    simplex_facet_rule_set<3>(fr, fpts, fwts, nfq);
    // k_tet_cell_tensors (GPU-tested on its own): the same routines into the SoA buffers
    for (int c = 0; c < E; ++c) {
        SimplexCell<3> cd;
        double X[4][3];
        for (int a = 0; a < 4; ++a) {
            const int v = cells[4 * (int64_t)c + a];
            for (int k = 0; k < 3; ++k) {
                X[a][k] = x[3 * (int64_t)v + k];
                cd.U[a][k] = sol[3 * (int64_t)v + k];
                cd.N[a][k] = un[3 * (int64_t)v + k];
                cd.H[a][k] = uh[3 * (int64_t)v + k];
            }
            cd.P[a] = sol[3 * (int64_t)n + v];
        }
        for (int k = 0; k < 3; ++k) cd.fbody[k] = g_f[k];
        cd.h = h[c];
        simplex_geometry<3>(cd, X);
        simplex_derive<3>(cd, g_par);
        double Fu[4][3], Fp[4];
        simplex_residual<3>(cd, g_par, g_r3.r[0], g_r3.r[1], Fu, Fp);
        for (int a = 0; a < 4; ++a) {
            for (int k = 0; k < 3; ++k) Fe[(int64_t)(a * 4 + k) * E + c] = Fu[a][k];
            Fe[(int64_t)(a * 4 + 3) * E + c] = Fp[a];
        }
        simplex_jacobian<3>(cd, g_par, g_r3.r[2], g_r3.r[3], g_r3.r[4], g_r3.r[5],
                            [&](int a, int bb, int ri, int ci, double v) {
                                Ae[(int64_t)((a * 4 + bb) * 16 + ri * 4 + ci) * E + c] = v;
                            });
    }
    // k_tet_facets
    for (int t = 0; t < m; ++t)
        tet_facet_item(t, E, n, fcells, fmask, co, fr, cells, x, h, sol, un, g_par, true, Ae, Fe);
    // k_gather_matrix3d
    for (int64_t s = 0; s < nnz_node; ++s)
        tet_gather_matrix_item(s, n, nnz_node, E, nrowptr, ncol, rowof, mseg_ptr, mseg_src, Ae, dofflag, dofmult, vals);
    // k_tet_lift_vector + k_tet_lift
    if (dofflag) {
        double* dvec = new double[4 * (int64_t)n];
        for (int64_t i = 0; i < 4 * (int64_t)n; ++i) dvec[i] = dofflag[i] ? (g[i] - sol[i]) : 0.0;
        for (int c = 0; c < E; ++c)
            if (cellflag[c]) tet_lift_item(c, E, n, cells, dvec, Ae, Fe);
        delete[] dvec;
    }
    // k_gather_vector3d
    for (int i = 0; i < n; ++i) tet_gather_vector_item(i, n, E, vseg_ptr, vseg_src, Fe, dofflag, sol, g, b);
    // k_spmv_node3d: 8 lanes per node, reduced in shuffle order (pairwise tree over the lanes)
    for (int i = 0; i < n; ++i) {
        double part[8][4];
        for (int lane = 0; lane < 8; ++lane) tet_spmv_item(i, lane, 8, n, nnz_node, nrowptr, ncol, vals, xv, part[lane]);
        for (int o = 4; o > 0; o >>= 1)
            for (int lane = 0; lane < o; ++lane)
                for (int k = 0; k < 4; ++k) part[lane][k] += part[lane + o][k];
        for (int k = 0; k < 4; ++k) y[tet_dof(n, i, k)] = part[0][k];
    }
    // k_tet_facet_flux + k_sum_serial
    double q = 0.0;
    for (int t = 0; t < m; ++t) q += tet_flux_item(t, n, fcells, fmask, cells, x, un);
    flux[0] = q;
}

// Emulated hemo_tet_pc_setup + hemo_tet_velocity_solve (k_tet_dinv, k_tet_a01_residual, k_tet_jacobi with
// the library's ping-pong order): zu after `sweeps` damped block-Jacobi sweeps on A00 zu = ru - A01 zp.
void txh_velocity_solve(int n, int64_t nnz_node, const int32_t* nrowptr, const int32_t* ncol, const int32_t* diagslot,
                        const double* vals, const double* ru, const double* zp, int sweeps, double omega, double* dinv,
                        double* tu, double* tmp, double* zu) {
    for (int i = 0; i < n; ++i) tet_dinv_item(i, n, nnz_node, nrowptr, diagslot, vals, dinv);
    for (int i = 0; i < n; ++i) tet_a01_residual_item(i, n, nnz_node, nrowptr, ncol, vals, zp, ru, tu);
    const double* zin = nullptr;
    for (int s = 0; s < sweeps; ++s) {
        double* zout = ((sweeps - 1 - s) & 1) ? tmp : zu;
        for (int i = 0; i < n; ++i) tet_jacobi_item(i, n, nnz_node, nrowptr, ncol, vals, dinv, omega, tu, zin, zout);
        zin = zout;
    }
}

// Emulated k_tet_selfp: Sp = A11 - A10 diag(A00)^-1 A01 on the distance-2 pattern (rowof2, col2)
void txh_selfp(int64_t nnz2, int64_t nnz_node, const int32_t* rowof2, const int32_t* col2, const int32_t* nrowptr,
               const int32_t* ncol, const int32_t* diagslot, const double* vals, double* out) {
    for (int64_t s = 0; s < nnz2; ++s) tet_selfp_item<double>(s, nnz_node, rowof2, col2, nrowptr, ncol, diagslot, vals, out);
}

// Emulated k_tet_l2_partial / k_tet_wss (post-processing on tetrahedra); wss: 3n, zero-initialised by the caller
double txh_l2(int E, int bs, const int32_t* cells, const double* x, const double* f) {
    double acc = 0.0;
    for (int c = 0; c < E; ++c) acc += tet_l2_item(c, bs, cells, x, f);
    return acc;
}

void txh_wss(int m, const int32_t* fcells, const int32_t* fmask, const int32_t* cells, const double* x, const double* sol,
             double mu, double* wss) {
    for (int t = 0; t < m; ++t)
        tet_wss_item(t, fcells, fmask, cells, x, sol, mu, [wss](int node, int k, double v) { wss[3 * (int64_t)node + k] += v; });
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Curl-curl / rotational formulation (csrc/curlcurl_element.cuh): residual with the rules of blocks 0 (F_u rows)
// and 1 (F_p rows), Jacobian rows with the rules of blocks 2 (velocity rows) and 4 (pressure rows).
// ---------------------------------------------------------------------------
#include "../../cfd_hemodynamic_b200/csrc/curlcurl_element.cuh"

template <int D>
static void run_curlcurl(Rules<D>& R, int E, int n, const int32_t* cells, const double* x, const double* h,
                         const double* sol, const double* un, double* Fe, double* Je) {
    constexpr int NV = D + 1, NL = NV * NV;
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
            }
            cd.P[a] = sol[D * (int64_t)n + v];
        }
        for (int k = 0; k < D; ++k) cd.fbody[k] = g_f[k];
        cd.h = h[c];
        simplex_geometry<D>(cd, X);
        double Fu[NV][D] = {}, Fp[NV] = {}, Fu2[NV][D] = {}, Fp2[NV] = {};
        double* J = Je + (int64_t)c * NL * NL;
        curlcurl_cell<D>(cd, g_par, R.r[0], true, false, false, Fu, Fp, nullptr);
        curlcurl_cell<D>(cd, g_par, R.r[1], false, true, false, Fu, Fp, nullptr);
        curlcurl_cell<D>(cd, g_par, R.r[2], true, false, true, Fu2, Fp2, J);
        curlcurl_cell<D>(cd, g_par, R.r[4], false, true, true, Fu2, Fp2, J);
        double* F = Fe + (int64_t)c * NL;
        for (int a = 0; a < NV; ++a) {
            for (int k = 0; k < D; ++k) F[a * D + k] = Fu[a][k];
            F[D * NV + a] = Fp[a];
        }
    }
}

extern "C" {

// Fe: [E][(D+1)^2] (velocity rows a*D + k, then pressure rows), Je: [E][(D+1)^2][(D+1)^2], zero-initialised
void cch_cells(int dim, int E, int n, const int32_t* cells, const double* x, const double* h, const double* sol,
               const double* un, double* Fe, double* Je) {
    if (dim == 2) run_curlcurl<2>(g_r2, E, n, cells, x, h, sol, un, Fe, Je);
    else run_curlcurl<3>(g_r3, E, n, cells, x, h, sol, un, Fe, Je);
}

}  // extern "C"

template <int D>
static void run_curlcurl_facets(int m, int n, const int32_t* pairs, const int32_t* cells, const double* x, const double* h,
                                const double* sol, const double* un, const hemo_facet_coef* co, const double* fpts,
                                const double* fwts, int nq, double* Fu, double* J) {
    constexpr int NV = D + 1;
    hemo_form_finalize(g_par);
    SimplexFacetRule<D> fr;
    simplex_facet_rule_set<D>(fr, fpts, fwts, nq);
    for (int t = 0; t < m; ++t) {
        const int c = pairs[2 * t], lf = pairs[2 * t + 1];
        SimplexCell<D> cd;
        double X[NV][D];
        for (int a = 0; a < NV; ++a) {
            const int v = cells[NV * (int64_t)c + a];
            for (int k = 0; k < D; ++k) {
                X[a][k] = x[D * (int64_t)v + k];
                cd.U[a][k] = sol[D * (int64_t)v + k];
                cd.N[a][k] = un[D * (int64_t)v + k];
            }
        }
        cd.h = h[c];
        simplex_geometry<D>(cd, X);
        double* fu = Fu + (int64_t)t * NV * D;
        double* jj = J + (int64_t)t * NV * NV * D * D;
        curlcurl_facet<D, true, true>(
            cd, g_par, *co, fr, lf, [&](int a, int k, double v) { fu[a * D + k] += v; },
            [&](int a, int b, int k, int l, double v) { jj[((a * NV + b) * D + k) * D + l] += v; });
    }
}

extern "C" {

// Fu: [m][NV][D], J: [m][NV][NV][D][D] (zero-initialised by the caller)
void cch_facets(int dim, int m, int n, const int32_t* pairs, const int32_t* cells, const double* x, const double* h,
                const double* sol, const double* un, const double* coef8, const double* fpts, const double* fwts, int nq,
                double* Fu, double* J) {
    hemo_facet_coef co = {coef8[0], coef8[1], coef8[2], coef8[3], coef8[4], coef8[5], coef8[6], coef8[7]};
    if (dim == 2) run_curlcurl_facets<2>(m, n, pairs, cells, x, h, sol, un, &co, fpts, fwts, nq, Fu, J);
    else run_curlcurl_facets<3>(m, n, pairs, cells, x, h, sol, un, &co, fpts, fwts, nq, Fu, J);
}

}  // extern "C"
