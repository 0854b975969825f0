// P1–P1 tetrahedron element-tensor kernel for sm_100a: the moment-factorised routines of
// simplex_element.cuh (D = 3) behind a C-ABI entry that fills the SoA element buffers
//     Ae[(a*4+b)*16 + ri*4+ci][E]  (ri/ci in u_x, u_y, u_z, p)   and   Fe[a*4+comp][E].
//
// Replaces the FFCx tetrahedron `tabulate_tensor` kernels of the forms at
// src/solvers/stabilized_schur.py:60-123 (reached through `mesh.topology.cell_name()`, e.g.
// src/scenarios/taylor_green.py:34), followed by the exterior-facet terms on triangular facets,
// the Dirichlet treatment (lifting, zeroed rows / columns, set_bc), the atomic-free gather into the
// reference's CSR layout and the matrix-vector product on that layout.  The per-thread bodies of
// everything after the cell kernel live in tet_items.cuh (host-checkable).  The 3-D multigrid
// preconditioner is the follow-up.
//
// One thread per cell.  The rule tables (up to 7^3 collapsed Gauss–Jacobi points per block form,
// 13.7 kB each) do not fit constant memory and live in global memory: every thread of a warp reads
// the same entry, so the loads are broadcasts served by L1.  Rules shared by several block forms
// (alias table from the host) are integrated once.
#include "hemo_internal.cuh"
#include "simplex_element.cuh"
#include "tet_items.cuh"

enum { TET_JAC_ALL = 0, TET_JAC_FLAGGED = 1, TET_JAC_NONE = 2 };

struct TetRules {
    SimplexRule<3> r[HEMO_NRULES];
    int alias[HEMO_NRULES];          // lowest block id with an identical rule
};

struct hemo_tet_state {
    TetRules* host = nullptr;
    TetRules* dev = nullptr;
    SimplexFacetRule<3> frule{};     // triangle rule of the exterior-facet integrals (kernel argument)
    bool have_frule = false;
    double* dinv = nullptr;          // 9n: inverse 3x3 node-diagonal blocks of A00 (block-Jacobi sweeps of the PC)
    int dinv_n = 0;
    bool have[HEMO_NRULES] = {false, false, false, false, false, false};
    bool dirty = true;
};

static hemo_tet_state* tet_state(hemo_ctx* ctx) {
    if (!ctx->tet) ctx->tet = new hemo_tet_state();
    return ctx->tet;
}

__global__ void __launch_bounds__(128)
k_tet_cell_tensors(int E, int n, const int32_t* __restrict__ cells, const double* __restrict__ x,
                   const double* __restrict__ h, const double* __restrict__ sol, const double* __restrict__ un,
                   const double* __restrict__ uh, HemoForm par, double f0, double f1, double f2,
                   const TetRules* __restrict__ rules, int jac_mode, const uint8_t* __restrict__ jac_cells,
                   double* __restrict__ Ae, double* __restrict__ Fe) {
    // jac_mode: TET_JAC_ALL = element Jacobian of every cell, TET_JAC_FLAGGED = only where jac_cells[c] != 0
    // (residual pass: the lifting needs it on Dirichlet-adjacent cells), TET_JAC_NONE = residual only
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= E) return;
    SimplexCell<3> cd;
    double X[4][3];
    const int4 vv = reinterpret_cast<const int4*>(cells)[c];
    const int v[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            X[a][k] = x[3 * (int64_t)v[a] + k];
            cd.U[a][k] = sol[3 * (int64_t)v[a] + k];
            cd.N[a][k] = un[3 * (int64_t)v[a] + k];
            cd.H[a][k] = uh[3 * (int64_t)v[a] + k];
        }
        cd.P[a] = sol[3 * (int64_t)n + v[a]];
    }
    cd.fbody[0] = f0; cd.fbody[1] = f1; cd.fbody[2] = f2;
    cd.h = h[c];
    simplex_geometry<3>(cd, X);
    simplex_derive<3>(cd, par);
    const int64_t stride = E;
    // moments of the distinct rules, block ids HEMO_Q_FU .. HEMO_Q_PP
    double T2[HEMO_NRULES > 0 ? 2 : 1][4][4];      // [0]: residual group scratch / J_uu, [1]: scratch
    double L0;
    // ---- residual: rules FU, FP
    {
        double T2p[4][4], L0p;
        simplex_moments<3>(cd, par, rules->r[HEMO_Q_FU], T2[0], L0);
        if (rules->alias[HEMO_Q_FP] == HEMO_Q_FU) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) T2p[a][b] = T2[0][a][b];
        } else {
            simplex_moments<3>(cd, par, rules->r[HEMO_Q_FP], T2p, L0p);
        }
        double T1p[4], T0;
        simplex_colsum<3>(T2p, T1p, T0);
        const SimplexRule<3>& ru = rules->r[HEMO_Q_FU];
        const SimplexRule<3>& rp = rules->r[HEMO_Q_FP];
        const double m0 = ru.m0 * cd.detJ;
        double pbar = 0.0;
#pragma unroll
        for (int b = 0; b < 4; ++b) pbar += ru.m1[b] * cd.P[b];
        pbar *= cd.detJ;
        for (int a = 0; a < 4; ++a) {
            double Wd[4];
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                double s = 0.0;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) s += T2[0][cc][d] * cd.s[cc][a];
                Wd[d] = s;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double val = 0.0;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) val += par.rho * cd.detJ * ru.m2[a][cc] * cd.A[cc][k];
                double sg = 0.0;
#pragma unroll
                for (int i = 0; i < 3; ++i) sg += cd.g[a][i] * par.mu * (cd.G[i][k] + cd.G[k][i]);
                val += m0 * sg - cd.g[a][k] * pbar;
#pragma unroll
                for (int d = 0; d < 4; ++d) val += Wd[d] * cd.R[d][k];
                val += L0 * par.rho * cd.divu * cd.g[a][k];
                Fe[(a * 4 + k) * stride + c] = val;
            }
            double acc = 0.0;
#pragma unroll
            for (int d = 0; d < 4; ++d)
                acc += T1p[d] * (cd.R[d][0] * cd.g[a][0] + cd.R[d][1] * cd.g[a][1] + cd.R[d][2] * cd.g[a][2]);
            Fe[(a * 4 + 3) * stride + c] = cd.detJ * rp.m1[a] * cd.divu + acc / par.rho;
        }
    }
    if (jac_mode == TET_JAC_NONE || (jac_mode == TET_JAC_FLAGGED && !jac_cells[c])) return;
    // ---- Jacobian: rules UU, UP, PU, PP (alias = lowest identical block id among all six)
    {
        double T1up[4], T1pu[4], T1[4], T0, T0pp = 0.0, Lt;
        auto col_of = [&](int blk, double T1out[4], double& T0out) {
            simplex_moments<3>(cd, par, rules->r[blk], T2[1], Lt);
            simplex_colsum<3>(T2[1], T1out, T0out);
        };
        col_of(HEMO_Q_UP, T1up, T0);
        if (rules->alias[HEMO_Q_PU] == rules->alias[HEMO_Q_UP]) {
#pragma unroll
            for (int d = 0; d < 4; ++d) T1pu[d] = T1up[d];
        } else {
            col_of(HEMO_Q_PU, T1pu, T0);
        }
        if (rules->alias[HEMO_Q_PP] == rules->alias[HEMO_Q_UP]) {
            T0pp = T1up[0] + T1up[1] + T1up[2] + T1up[3];
        } else if (rules->alias[HEMO_Q_PP] == rules->alias[HEMO_Q_PU]) {
            T0pp = T1pu[0] + T1pu[1] + T1pu[2] + T1pu[3];
        } else {
            col_of(HEMO_Q_PP, T1, T0pp);
        }
        if (rules->alias[HEMO_Q_UU] != HEMO_Q_FU) simplex_moments<3>(cd, par, rules->r[HEMO_Q_UU], T2[0], L0);
        // (alias FU: T2[0] / L0 still hold the moments of the identical F_u rule)
        double* out = Ae + c;
        simplex_jacobian_from_moments<3>(cd, par, T2[0], L0, T1up, T1pu, T0pp, rules->r[HEMO_Q_UU].m0,
                                         rules->r[HEMO_Q_UU].m2, rules->r[HEMO_Q_UP].m1, rules->r[HEMO_Q_PU].m1,
                                         [&](int a, int b, int ri, int ci, double val) {
                                             out[((a * 4 + b) * 16 + ri * 4 + ci) * stride] = val;
                                         });
    }
}

extern "C" int hemo_tet_set_quadrature(hemo_ctx* ctx, int block, const double* pts, const double* wts, int nq) {
    if (!ctx || block < 0 || block >= HEMO_NRULES || !pts || !wts || nq <= 0) return HEMO_EINVAL;
    if (nq > HEMO_SIMPLEX_MAXQ) HEMO_FAIL(ctx, HEMO_EINVAL, "too many quadrature points for a tetrahedron rule");
    hemo_tet_state* st = tet_state(ctx);
    if (!st->host) {
        st->host = (TetRules*)calloc(1, sizeof(TetRules));
        if (!st->host) HEMO_FAIL(ctx, HEMO_EINVAL, "out of host memory");
    }
    simplex_rule_set<3>(st->host->r[block], pts, wts, nq);
    st->have[block] = true;
    for (int b = 0; b < HEMO_NRULES; ++b) {
        if (!st->have[b]) continue;
        st->host->alias[b] = b;
        for (int a = 0; a < b; ++a) {
            if (!st->have[a] || st->host->r[a].nq != st->host->r[b].nq) continue;
            bool same = true;
            for (int q = 0; q < st->host->r[b].nq && same; ++q) {
                same = st->host->r[a].w[q] == st->host->r[b].w[q];
                for (int k = 0; k < 4 && same; ++k) same = st->host->r[a].phi[q][k] == st->host->r[b].phi[q][k];
            }
            if (same) { st->host->alias[b] = st->host->alias[a]; break; }
        }
    }
    st->dirty = true;
    return 0;
}

static int tet_launch_cells(hemo_ctx* ctx, int n_nodes, int n_cells, const double* x_dev, const int32_t* cells_dev,
                            const double* h_dev, const double* sol_dev, const double* un_dev, const double* uh_dev,
                            const double* f3_host, int jac_mode, const uint8_t* jac_cells, double* Ae_dev,
                            double* Fe_dev) {
    if (!ctx || n_nodes <= 0 || n_cells <= 0 || !x_dev || !cells_dev || !h_dev || !sol_dev || !un_dev || !f3_host ||
        !Ae_dev || !Fe_dev)
        return HEMO_EINVAL;
    if (!ctx->have_par) HEMO_FAIL(ctx, HEMO_ESTATE, "hemo_set_params not called");
    hemo_tet_state* st = tet_state(ctx);
    for (int r = 0; r < HEMO_NRULES; ++r)
        if (!st->have[r]) HEMO_FAIL(ctx, HEMO_ESTATE, "tetrahedron quadrature rule missing for a block form");
    if (st->dirty) {
        if (!st->dev) HEMO_CHECK_CUDA(ctx, cudaMalloc((void**)&st->dev, sizeof(TetRules)));
        HEMO_CHECK_CUDA(ctx, cudaMemcpyAsync(st->dev, st->host, sizeof(TetRules), cudaMemcpyHostToDevice, ctx->stream));
        st->dirty = false;
    }
    hemo_form_finalize(ctx->par);
    k_tet_cell_tensors<<<hemo_grid(n_cells, 128), 128, 0, ctx->stream>>>(
        n_cells, n_nodes, cells_dev, x_dev, h_dev, sol_dev, un_dev, uh_dev ? uh_dev : un_dev, ctx->par, f3_host[0],
        f3_host[1], f3_host[2], st->dev, jac_mode, jac_cells, Ae_dev, Fe_dev);
    HEMO_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int hemo_tet_element_tensors(hemo_ctx* ctx, int n_nodes, int n_cells, const double* x_dev,
                                        const int32_t* cells_dev, const double* h_dev, const double* sol_dev,
                                        const double* un_dev, const double* uh_dev, const double* f3_host,
                                        double* Ae_dev, double* Fe_dev) {
    return tet_launch_cells(ctx, n_nodes, n_cells, x_dev, cells_dev, h_dev, sol_dev, un_dev, uh_dev, f3_host, TET_JAC_ALL,
                            nullptr, Ae_dev, Fe_dev);
}

// ---------------------------------------------------------------------------
// 3-D assembly into the CSR the reference's create_matrix_block builds for P1-P1 tetrahedra
// (src/solvers/stabilized_schur.py:191-193): global vector [u interleaved (3n) | p (n)], rows in
// that order, columns ascending; layout formulas in tet_items.cuh.
// Same atomic-free gather as the 2-D path (fixed summation order), 16 scalars per node pair.
// ---------------------------------------------------------------------------
int hemo_ensure_elem(hemo_ctx* ctx, size_t ae_count, size_t fe_count);

__global__ void k_pattern3d(int n, int64_t nnz_node, const int32_t* __restrict__ nrowptr,
                            const int32_t* __restrict__ ncol, int64_t* __restrict__ rowptr,
                            int32_t* __restrict__ colind) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { rowptr[4 * (int64_t)n] = 16 * nnz_node; return; }
    const int r0 = nrowptr[i];
    const int deg = nrowptr[i + 1] - r0;
    int64_t rs[4];
    for (int k = 0; k < 3; ++k) { rs[k] = 12 * (int64_t)r0 + 4 * (int64_t)k * deg; rowptr[3 * (int64_t)i + k] = rs[k]; }
    rs[3] = 12 * nnz_node + 4 * (int64_t)r0;
    rowptr[3 * (int64_t)n + i] = rs[3];
    for (int t = 0; t < deg; ++t) {
        const int j = ncol[r0 + t];
        for (int k = 0; k < 4; ++k) {
            colind[rs[k] + 3 * t] = 3 * j; colind[rs[k] + 3 * t + 1] = 3 * j + 1; colind[rs[k] + 3 * t + 2] = 3 * j + 2;
            colind[rs[k] + 3 * deg + t] = 3 * n + j;
        }
    }
}

__global__ void __launch_bounds__(128)
k_tet_facets(int m, int64_t E, int n, const int32_t* __restrict__ fcells, const int32_t* __restrict__ fmask,
             hemo_facet_coef co, SimplexFacetRule<3> fr, const int32_t* __restrict__ cells,
             const double* __restrict__ x, const double* __restrict__ h, const double* __restrict__ sol,
             const double* __restrict__ un, HemoForm par, int jac_mode, const uint8_t* __restrict__ jac_cells,
             double* __restrict__ Ae, double* __restrict__ Fe) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    // the derivative goes into Ae exactly where the cell kernel of this pass wrote Ae
    const bool want_jac = jac_mode == TET_JAC_ALL || (jac_mode == TET_JAC_FLAGGED && jac_cells[fcells[t]]);
    tet_facet_item(t, E, n, fcells, fmask, co, fr, cells, x, h, sol, un, par, want_jac, Ae, Fe);
}

__global__ void __launch_bounds__(128)
k_tet_facet_flux(int m, int n, const int32_t* __restrict__ fcells, const int32_t* __restrict__ fmask,
                 const int32_t* __restrict__ cells, const double* __restrict__ x, const double* __restrict__ un,
                 double* __restrict__ partial) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    partial[t] = tet_flux_item(t, n, fcells, fmask, cells, x, un);
}

__global__ void __launch_bounds__(128)
k_tet_lift(int64_t E, int n, const int32_t* __restrict__ cells, const uint8_t* __restrict__ cellflag,
           const double* __restrict__ dvec, const double* __restrict__ Ae, double* __restrict__ Fe) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= E || !cellflag[c]) return;
    tet_lift_item(c, E, n, cells, dvec, Ae, Fe);
}

__global__ void k_tet_lift_vector(int64_t N, const uint8_t* __restrict__ dofflag, const double* __restrict__ x,
                                  const double* __restrict__ g, double* __restrict__ d) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    d[i] = dofflag[i] ? (g[i] - x[i]) : 0.0;
}

__global__ void __launch_bounds__(256)
k_gather_matrix3d(int n, int64_t nnz_node, int64_t E, const int32_t* __restrict__ nrowptr,
                  const int32_t* __restrict__ ncol, const int32_t* __restrict__ rowof,
                  const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ seg_src,
                  const double* __restrict__ Ae, const uint8_t* __restrict__ dofflag,
                  const double* __restrict__ dofmult, double* __restrict__ vals) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= nnz_node) return;
    tet_gather_matrix_item(s, n, nnz_node, E, nrowptr, ncol, rowof, seg_ptr, seg_src, Ae, dofflag, dofmult, vals);
}

__global__ void __launch_bounds__(256)
k_gather_vector3d(int n, int64_t E, const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ seg_src,
                  const double* __restrict__ Fe, const uint8_t* __restrict__ dofflag, const double* __restrict__ xk,
                  const double* __restrict__ g, double* __restrict__ b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    tet_gather_vector_item(i, n, E, seg_ptr, seg_src, Fe, dofflag, xk, g, b);
}

// y = J x on the 3-D CSR layout: 8 lanes per node (4 rows each), shuffle reduction, lane 0 stores.
// HBM-bound: 16 values (128 B) + one column index per node pair, four x entries per neighbour.
__global__ void __launch_bounds__(256)
k_spmv_node3d(int n, int64_t nnz_node, const int32_t* __restrict__ nrowptr, const int32_t* __restrict__ ncol,
              const double* __restrict__ vals, const double* __restrict__ xv, double* __restrict__ y) {
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = gt >> 3;
    const int lane = gt & 7;
    const bool ok = i < n;            // no early return: full-mask shuffles below
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (ok) tet_spmv_item(i, lane, 8, n, nnz_node, nrowptr, ncol, vals, xv, acc);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], o, 8);
    }
    if (ok && lane == 0) {
        y[3 * (int64_t)i] = acc[0]; y[3 * (int64_t)i + 1] = acc[1]; y[3 * (int64_t)i + 2] = acc[2];
        y[3 * (int64_t)n + i] = acc[3];
    }
}

int hemo_tet_pattern(hemo_ctx* ctx, int64_t* rowptr_dev, int32_t* colind_dev) {
    k_pattern3d<<<hemo_grid(ctx->n + 1, 256), 256, 0, ctx->stream>>>(ctx->n, ctx->nnz_node, ctx->nrowptr, ctx->ncol,
                                                                     rowptr_dev, colind_dev);
    HEMO_LAUNCH_CHECK(ctx);
    return 0;
}

int hemo_tet_set_facet_quadrature(hemo_ctx* ctx, const double* pts, const double* wts, int nq) {
    if (nq > HEMO_SIMPLEX_MAXFQ) HEMO_FAIL(ctx, HEMO_EINVAL, "too many points for a triangular facet rule");
    hemo_tet_state* st = tet_state(ctx);
    simplex_facet_rule_set<3>(st->frule, pts, wts, nq);
    st->have_frule = true;
    return 0;
}

static bool tet_facet_set_active(const HemoFacetSet& fs) {
    const hemo_facet_coef& c = fs.coef;
    return fs.m > 0 && (c.a_p != 0.0 || c.pconst != 0.0 || c.a_g != 0.0 || c.a_s != 0.0 || c.a_n != 0.0 || c.a_b != 0.0);
}

int hemo_cc_cells(hemo_ctx* ctx, const double* x_dev, const double* un_dev, int jac_mode);   // assembly_curlcurl.cu

// host-side rule tables for the curl-curl kernels (same rules, FFCx-style evaluation)
int hemo_tet_get_rules(hemo_ctx* ctx, const SimplexRule<3>** rules, const int** alias, const SimplexFacetRule<3>** frule) {
    hemo_tet_state* st = tet_state(ctx);
    for (int r = 0; r < HEMO_NRULES; ++r)
        if (!st->have[r]) HEMO_FAIL(ctx, HEMO_ESTATE, "tetrahedron quadrature rule missing for a block form");
    *rules = st->host->r;
    *alias = st->host->alias;
    *frule = st->have_frule ? &st->frule : nullptr;
    return 0;
}

// cell tensors (residual, and the Jacobian where jac_mode asks for it), then the facet terms of every active set
static int tet_cells(hemo_ctx* ctx, const double* x_dev, const double* un_dev, int jac_mode) {
    if (!ctx->cells || !ctx->nrowptr) HEMO_FAIL(ctx, HEMO_ESTATE, "mesh / node graph not set");
    int rc = hemo_ensure_elem(ctx, (size_t)256 * ctx->E, (size_t)16 * ctx->E);
    if (rc) return rc;
    if (ctx->formulation == HEMO_FORM_CURLCURL) return hemo_cc_cells(ctx, x_dev, un_dev, jac_mode);
    const double f3[3] = {ctx->par.f[0], ctx->par.f[1], ctx->fz};
    if (!ctx->have_par) HEMO_FAIL(ctx, HEMO_ESTATE, "hemo_set_params not called");
    if ((rc = tet_launch_cells(ctx, ctx->n, ctx->E, ctx->x, ctx->cells, ctx->h, x_dev, un_dev, ctx->uh, f3, jac_mode,
                               ctx->cellflag, ctx->Ae, ctx->Fe)))
        return rc;
    hemo_tet_state* st = tet_state(ctx);
    for (int s = 0; s < HEMO_MAX_FACET_SETS; ++s) {
        const HemoFacetSet& fs = ctx->fsets[s];
        if (!tet_facet_set_active(fs)) continue;
        if (!st->have_frule) HEMO_FAIL(ctx, HEMO_ESTATE, "hemo_set_facet_quadrature not called (triangle rule)");
        k_tet_facets<<<hemo_grid(fs.m, 128), 128, 0, ctx->stream>>>(fs.m, ctx->E, ctx->n, fs.cells, fs.mask, fs.coef,
                                                                    st->frule, ctx->cells, ctx->x, ctx->h, x_dev, un_dev,
                                                                    ctx->par, jac_mode, ctx->cellflag, ctx->Ae, ctx->Fe);
        HEMO_LAUNCH_CHECK(ctx);
    }
    return 0;
}

int hemo_tet_assemble_jacobian(hemo_ctx* ctx, const double* x_dev, const double* un_dev, double* vals_dev) {
    int rc = tet_cells(ctx, x_dev, un_dev, TET_JAC_ALL);
    if (rc) return rc;
    k_gather_matrix3d<<<hemo_grid(ctx->nnz_node, 256), 256, 0, ctx->stream>>>(
        ctx->n, ctx->nnz_node, ctx->E, ctx->nrowptr, ctx->ncol, ctx->rowof, ctx->mseg_ptr, ctx->mseg_src, ctx->Ae,
        ctx->have_bc ? ctx->dofflag : nullptr, ctx->dofmult, vals_dev);
    HEMO_LAUNCH_CHECK(ctx);
    return 0;
}

int hemo_tet_assemble_residual(hemo_ctx* ctx, const double* x_dev, const double* un_dev, const double* g_dev,
                               double* b_dev) {
    if (ctx->have_bc && !g_dev) return HEMO_EINVAL;
    int rc = tet_cells(ctx, x_dev, un_dev, ctx->have_bc ? TET_JAC_FLAGGED : TET_JAC_NONE);
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    if (ctx->have_bc) {
        const int64_t N = 4 * (int64_t)ctx->n;
        k_tet_lift_vector<<<hemo_grid(N, 256), 256, 0, st>>>(N, ctx->dofflag, x_dev, g_dev, ctx->dvec);
        HEMO_LAUNCH_CHECK(ctx);
        k_tet_lift<<<hemo_grid(ctx->E, 128), 128, 0, st>>>(ctx->E, ctx->n, ctx->cells, ctx->cellflag, ctx->dvec, ctx->Ae,
                                                           ctx->Fe);
        HEMO_LAUNCH_CHECK(ctx);
    }
    k_gather_vector3d<<<hemo_grid(ctx->n, 256), 256, 0, st>>>(ctx->n, ctx->E, ctx->vseg_ptr, ctx->vseg_src, ctx->Fe,
                                                              ctx->have_bc ? ctx->dofflag : nullptr, x_dev, g_dev, b_dev);
    HEMO_LAUNCH_CHECK(ctx);
    return 0;
}

int hemo_tet_spmv(hemo_ctx* ctx, const double* vals_dev, const double* x_dev, double* y_dev) {
    k_spmv_node3d<<<hemo_grid(8 * (int64_t)ctx->n, 256), 256, 0, ctx->stream>>>(ctx->n, ctx->nnz_node, ctx->nrowptr,
                                                                                ctx->ncol, vals_dev, x_dev, y_dev);
    HEMO_LAUNCH_CHECK(ctx);
    return 0;
}

int hemo_tet_facet_flux(hemo_ctx* ctx, const HemoFacetSet& fs, const double* un_dev, double* partial) {
    k_tet_facet_flux<<<hemo_grid(fs.m, 128), 128, 0, ctx->stream>>>(fs.m, ctx->n, fs.cells, fs.mask, ctx->cells, ctx->x,
                                                                    un_dev, partial);
    HEMO_LAUNCH_CHECK(ctx);
    return 0;
}

__global__ void k_tet_cell_laplace(int64_t E, const int32_t* __restrict__ cells, const double* __restrict__ x,
                                   double* __restrict__ Ke, double* __restrict__ Me) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= E) return;
    tet_laplace_item(c, E, cells, x, Ke, Me);
}

// element stiffness / lumped mass into Ae[0..16E) / Fe[0..4E) (gathered by the generic scalar kernels)
int hemo_tet_laplace_mass(hemo_ctx* ctx) {
    k_tet_cell_laplace<<<hemo_grid(ctx->E, 256), 256, 0, ctx->stream>>>(ctx->E, ctx->cells, ctx->x, ctx->Ae, ctx->Fe);
    HEMO_LAUNCH_CHECK(ctx);
    return 0;
}

// ---------------------------------------------------------------------------
// 3-D preconditioner kernels (bodies in tet_items.cuh)
// ---------------------------------------------------------------------------
__global__ void k_tet_dinv(int n, int64_t nnz_node, const int32_t* __restrict__ nrowptr,
                           const int32_t* __restrict__ diagslot, const double* __restrict__ vals,
                           double* __restrict__ dinv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    tet_dinv_item(i, n, nnz_node, nrowptr, diagslot, vals, dinv);
}

__global__ void k_tet_a01_residual(int n, int64_t nnz_node, const int32_t* __restrict__ nrowptr,
                                   const int32_t* __restrict__ ncol, const double* __restrict__ vals,
                                   const double* __restrict__ zp, const double* __restrict__ ru,
                                   double* __restrict__ tu) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    tet_a01_residual_item(i, n, nnz_node, nrowptr, ncol, vals, zp, ru, tu);
}

__global__ void k_tet_jacobi(int n, int64_t nnz_node, const int32_t* __restrict__ nrowptr,
                             const int32_t* __restrict__ ncol, const double* __restrict__ vals,
                             const double* __restrict__ dinv, double omega, const double* __restrict__ tu,
                             const double* zin, double* zout) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    tet_jacobi_item(i, n, nnz_node, nrowptr, ncol, vals, dinv, omega, tu, zin, zout);
}

__global__ void __launch_bounds__(256)
k_tet_selfp(int64_t nnz2, int64_t nnz_node, const int32_t* __restrict__ rowof2, const int32_t* __restrict__ col2,
            const int32_t* __restrict__ nrowptr, const int32_t* __restrict__ ncol, const int32_t* __restrict__ diagslot,
            const double* __restrict__ vals, areal* __restrict__ out) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= nnz2) return;
    tet_selfp_item<areal>(s, nnz_node, rowof2, col2, nrowptr, ncol, diagslot, vals, out);
}

int hemo_tet_selfp(hemo_ctx* ctx, const double* vals_dev) {
    HemoAmg& amg = ctx->amg[1];
    k_tet_selfp<<<hemo_grid(amg.fine_nnz, 256), 256, 0, ctx->stream>>>(amg.fine_nnz, ctx->nnz_node, amg.fine_rowof,
                                                                       amg.fine_col, ctx->nrowptr, ctx->ncol, ctx->diagslot,
                                                                       vals_dev, amg.op[0].val);
    HEMO_LAUNCH_CHECK(ctx);
    return 0;
}

int hemo_tet_pc_setup(hemo_ctx* ctx, const double* vals_dev) {
    hemo_tet_state* st = tet_state(ctx);
    int rc;
    if (!st->dinv || st->dinv_n != ctx->n) {
        if ((rc = hemo_alloc(ctx, &st->dinv, (size_t)9 * ctx->n))) return rc;
        st->dinv_n = ctx->n;
    }
    k_tet_dinv<<<hemo_grid(ctx->n, 256), 256, 0, ctx->stream>>>(ctx->n, ctx->nnz_node, ctx->nrowptr, ctx->diagslot,
                                                                vals_dev, st->dinv);
    HEMO_LAUNCH_CHECK(ctx);
    return 0;
}

// t_u = r_u - A01 z_p, then `sweeps` damped block-Jacobi sweeps on A00 z_u = t_u from z_u = 0.
// tmp: 3n scratch (ping-pong partner of zu); the last sweep always lands in zu.
int hemo_tet_velocity_solve(hemo_ctx* ctx, const double* vals_dev, const double* ru, const double* zp, double* tu,
                            double* tmp, double* zu, int sweeps, double omega) {
    hemo_tet_state* st = tet_state(ctx);
    if (!st->dinv) HEMO_FAIL(ctx, HEMO_ESTATE, "hemo_pc_setup not called");
    const int n = ctx->n;
    const int grid = hemo_grid(n, 256);
    k_tet_a01_residual<<<grid, 256, 0, ctx->stream>>>(n, ctx->nnz_node, ctx->nrowptr, ctx->ncol, vals_dev, zp, ru, tu);
    HEMO_LAUNCH_CHECK(ctx);
    if (sweeps < 1) sweeps = 1;
    // sweep s writes buf[(sweeps - 1 - s) & 1] with buf[0] = zu, so that the final sweep writes zu
    const double* zin = nullptr;
    for (int s = 0; s < sweeps; ++s) {
        double* zout = ((sweeps - 1 - s) & 1) ? tmp : zu;
        k_tet_jacobi<<<grid, 256, 0, ctx->stream>>>(n, ctx->nnz_node, ctx->nrowptr, ctx->ncol, vals_dev, st->dinv, omega, tu,
                                                    zin, zout);
        HEMO_LAUNCH_CHECK(ctx);
        zin = zout;
    }
    return 0;
}

void hemo_tet_free(hemo_ctx* ctx) {
    if (!ctx->tet) return;
    free(ctx->tet->host);
    cudaFree(ctx->tet->dev);
    cudaFree(ctx->tet->dinv);
    delete ctx->tet;
    ctx->tet = nullptr;
}
