// P2-P2 triangle cell and exterior-facet kernels for sm_100a: the element routines of p2_element.cuh wrapped in
// load / call / store kernels that write the generic SoA element buffers (nv = 6: Ae[(a*6+b)*9 + ri*3+ci][E],
// Fe[a*3+comp][E]) the atomic-free gather kernels of assembly.cu read.
//
// Replaces the FFCx kernels behind assemble_matrix_block / assemble_vector_block for `p_grade = 2`
// (reference src/solvers/stabilized_schur_pressure_backflow.py:71,102-161,224-226; stabilized_schur_backflow.py:63,85-149).
//
// Work decomposition: the 18 x 18 element tensor at the reference quadrature (degree 20: 121 collapsed Gauss-Jacobi
// points, 79 with the Basix Xiao-Gimbutas rule) is FP64-pipe bound by two orders of magnitude over its 2.6 kB of output.
// One thread per (cell, test node): blockIdx.y is the test node, so the 54 stores of a thread are coalesced 256-byte
// lines across the warp; the point set-up (basis, state, tau) is recomputed by the six work items of a cell rather
// than exchanged through shared memory.  Nothing here is a dense contraction: per point the update is a rank-1-like
// sum of five products per entry with coefficients that change from point to point (tau, R, G), see DESIGN.md §4e.
#include "hemo_internal.cuh"
#include "p2_element.cuh"

__constant__ HemoP2Rule c_p2rules[HEMO_NRULES];
__constant__ HemoFacetRule c_p2frule;
__constant__ HemoForm c_p2par;

__device__ __forceinline__ void p2_load(P2Cell& cd, int c, const int32_t* __restrict__ cells, const double* __restrict__ x,
                                        const double* __restrict__ h, const double* __restrict__ sol,
                                        const double* __restrict__ un, const double* __restrict__ uh, int n, int v[6]) {
    const int2* cp = reinterpret_cast<const int2*>(cells + 6 * (int64_t)c);
    const int2 v01 = cp[0], v23 = cp[1], v45 = cp[2];
    v[0] = v01.x; v[1] = v01.y; v[2] = v23.x; v[3] = v23.y; v[4] = v45.x; v[5] = v45.y;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        if (a < 3) {
            const double2 xv = reinterpret_cast<const double2*>(x)[v[a]];
            cd.X[a][0] = xv.x; cd.X[a][1] = xv.y;
        }
        const double2 uv = reinterpret_cast<const double2*>(sol)[v[a]];
        cd.U[a][0] = uv.x; cd.U[a][1] = uv.y;
        const double2 nv = reinterpret_cast<const double2*>(un)[v[a]];
        cd.N[a][0] = nv.x; cd.N[a][1] = nv.y;
        const double2 hv = reinterpret_cast<const double2*>(uh)[v[a]];
        cd.H[a][0] = hv.x; cd.H[a][1] = hv.y;
        cd.P[a] = sol[2 * (int64_t)n + v[a]];
    }
    cd.h = h[c];
    p2_prepare(cd, c_p2par);
}

// one pass (rule `rule`, block forms MASK) of the Jacobian: blockIdx.y = test node
template <int MASK>
__global__ void __launch_bounds__(128)
k_p2_cell_jacobian(int E, int n, int rule, const int32_t* __restrict__ cells, const double* __restrict__ x,
                   const double* __restrict__ h, const double* __restrict__ sol, const double* __restrict__ un,
                   const double* __restrict__ uh, double* __restrict__ Ae) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= E) return;
    const int a = blockIdx.y;               // test node of this work item
    P2Cell cd;
    int v[6];
    p2_load(cd, c, cells, x, h, sol, un, uh, n, v);
    double* out = Ae + c;
    const int64_t stride = E;
    p2_jac_pass_node<MASK>(cd, c_p2par, c_p2rules[rule], a, [&](int slot, double val) { out[slot * stride] = val; });
}

// lifting of a boundary-adjacent cell: Fe += Ae (g - x), out of line (it recomputes the Jacobian rows)
__device__ __noinline__ void p2_lift_device(int c, int E, int n, const int32_t* __restrict__ cells, const double* __restrict__ x,
                                            const double* __restrict__ h, const double* __restrict__ sol,
                                            const double* __restrict__ un, const double* __restrict__ uh,
                                            const double* __restrict__ dvec, double* __restrict__ Fe) {
    P2Cell cd;
    int v[6];
    p2_load(cd, c, cells, x, h, sol, un, uh, n, v);
    double dl[6][3];
    bool any = false;
    for (int b = 0; b < 6; ++b) {
        dl[b][0] = dvec[2 * (int64_t)v[b]];
        dl[b][1] = dvec[2 * (int64_t)v[b] + 1];
        dl[b][2] = dvec[2 * (int64_t)n + v[b]];
        any = any || dl[b][0] != 0.0 || dl[b][1] != 0.0 || dl[b][2] != 0.0;
    }
    if (!any) return;
    const int64_t stride = E;
    for (int a = 0; a < 6; ++a) {
        double F[3] = {0.0, 0.0, 0.0};
        auto acc = [&](int slot, double val) {
            const int r = slot % 9, b = (slot / 9) % 6;
            F[r / 3] += val * dl[b][r % 3];
        };
        // block by block (boundary-adjacent cells only: register pressure matters more than the repeated point set-up)
        p2_jac_pass_node<P2_UU>(cd, c_p2par, c_p2rules[HEMO_Q_UU], a, acc);
        p2_jac_pass_node<P2_UP>(cd, c_p2par, c_p2rules[HEMO_Q_UP], a, acc);
        p2_jac_pass_node<P2_PU>(cd, c_p2par, c_p2rules[HEMO_Q_PU], a, acc);
        p2_jac_pass_node<P2_PP>(cd, c_p2par, c_p2rules[HEMO_Q_PP], a, acc);
        Fe[(a * 3 + 0) * stride + c] += F[0];
        Fe[(a * 3 + 1) * stride + c] += F[1];
        Fe[(a * 3 + 2) * stride + c] += F[2];
    }
}

__global__ void __launch_bounds__(128)
k_p2_cell_residual(int E, int n, const int32_t* __restrict__ cells, const double* __restrict__ x, const double* __restrict__ h,
                   const double* __restrict__ sol, const double* __restrict__ un, const double* __restrict__ uh,
                   const uint8_t* __restrict__ cellflag, const double* __restrict__ dvec, double* __restrict__ Fe) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= E) return;
    {
        P2Cell cd;
        int v[6];
        p2_load(cd, c, cells, x, h, sol, un, uh, n, v);
        double Fu[6][2], Fp[6];
        p2_cell_residual(cd, c_p2par, c_p2rules, Fu, Fp);
        const int64_t stride = E;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            Fe[(a * 3 + 0) * stride + c] = Fu[a][0];
            Fe[(a * 3 + 1) * stride + c] = Fu[a][1];
            Fe[(a * 3 + 2) * stride + c] = Fp[a];
        }
    }
    if (cellflag != nullptr && cellflag[c]) p2_lift_device(c, E, n, cells, x, h, sol, un, uh, dvec, Fe);
}

// One thread per boundary cell of a tagged set; mode 0: residual (+ lifting) into Fe, mode 1: Jacobian into Ae.
template <int MODE>
__global__ void __launch_bounds__(64)
k_p2_facets(int m, int E, int n, const int32_t* __restrict__ fcells, const int32_t* __restrict__ fmask, hemo_facet_coef co,
            const int32_t* __restrict__ cells, const double* __restrict__ x, const double* __restrict__ h,
            const double* __restrict__ sol, const double* __restrict__ un, const double* __restrict__ uh,
            const uint8_t* __restrict__ cellflag, const double* __restrict__ dvec, double* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const int c = fcells[t];
    const int mask = fmask[t];
    P2Cell cd;
    int v[6];
    p2_load(cd, c, cells, x, h, sol, un, uh, n, v);
    const int64_t stride = E;
    if (MODE == 1) {
        for (int b = 0; b < 6; ++b)
            for (int ci = 0; ci < 3; ++ci) {
                double col[6][2];
                p2_facet_column(cd, c_p2par, c_p2frule, co, mask, b, ci, col);
                for (int a = 0; a < 6; ++a)
                    for (int ri = 0; ri < 2; ++ri) out[((a * 6 + b) * 9 + ri * 3 + ci) * stride + c] += col[a][ri];
            }
    } else {
        double Fu[6][2];
        p2_facet_residual(cd, c_p2par, c_p2frule, co, mask, Fu);
        if (cellflag != nullptr && cellflag[c]) {
            // lifting: the facet operator is linear, apply it to (theta d_u, d_p)
            double Um[6][2], Pv[6], L[6][2];
            bool any = false;
            for (int b = 0; b < 6; ++b) {
                Um[b][0] = c_p2par.theta * dvec[2 * (int64_t)v[b]];
                Um[b][1] = c_p2par.theta * dvec[2 * (int64_t)v[b] + 1];
                Pv[b] = dvec[2 * (int64_t)n + v[b]];
                any = any || Um[b][0] != 0.0 || Um[b][1] != 0.0 || Pv[b] != 0.0;
            }
            if (any) {
                p2_facet_eval(cd, c_p2par, c_p2frule, co, mask, Um, Pv, false, L);
                for (int a = 0; a < 6; ++a) { Fu[a][0] += L[a][0]; Fu[a][1] += L[a][1]; }
            }
        }
        for (int a = 0; a < 6; ++a) {
            out[(a * 3 + 0) * stride + c] += Fu[a][0];
            out[(a * 3 + 1) * stride + c] += Fu[a][1];
        }
    }
}

__global__ void k_p2_facet_flux(int m, const int32_t* __restrict__ fcells, const int32_t* __restrict__ fmask,
                                const int32_t* __restrict__ cells, const double* __restrict__ x, const double* __restrict__ un,
                                double* __restrict__ partial) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const int c = fcells[t];
    P2Cell cd;
    for (int a = 0; a < 6; ++a) {
        const int v = cells[6 * (int64_t)c + a];
        if (a < 3) { cd.X[a][0] = x[2 * (int64_t)v]; cd.X[a][1] = x[2 * (int64_t)v + 1]; }
        cd.N[a][0] = un[2 * (int64_t)v]; cd.N[a][1] = un[2 * (int64_t)v + 1];
    }
    partial[t] = p2_cell_flux(cd, c_p2frule, fmask[t]);
}

__global__ void k_p2_cell_laplace(int E, const int32_t* __restrict__ cells, const double* __restrict__ x,
                                  double* __restrict__ Ke /*[36][E]*/, double* __restrict__ Me /*[6][E]*/) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= E) return;
    P2Cell cd;
    for (int a = 0; a < 3; ++a) {
        const int v = cells[6 * (int64_t)c + a];
        cd.X[a][0] = x[2 * (int64_t)v]; cd.X[a][1] = x[2 * (int64_t)v + 1];
    }
    const double J00 = cd.X[1][0] - cd.X[0][0], J01 = cd.X[2][0] - cd.X[0][0];
    const double J10 = cd.X[1][1] - cd.X[0][1], J11 = cd.X[2][1] - cd.X[0][1];
    const double det = J00 * J11 - J01 * J10, id = 1.0 / det;
    cd.K[0][0] = J11 * id; cd.K[0][1] = -J01 * id; cd.K[1][0] = -J10 * id; cd.K[1][1] = J00 * id;
    cd.adet = fabs(det);
    double K[6][6], M[6];
    p2_cell_laplace_mass(cd, K, M);
    for (int a = 0; a < 6; ++a) {
        for (int b = 0; b < 6; ++b) Ke[(int64_t)(a * 6 + b) * E + c] = K[a][b];
        Me[(int64_t)a * E + c] = M[a];
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
int hemo_p2_set_quadrature(hemo_ctx* ctx, int block, const double* pts, const double* wts, int nq) {
    if (nq > HEMO_MAXQ_P2) HEMO_FAIL(ctx, HEMO_EINVAL, "too many quadrature points for a P2 triangle rule (max 128)");
    if (!ctx->p2rules) {
        ctx->p2rules = calloc(HEMO_NRULES, sizeof(HemoP2Rule));
        if (!ctx->p2rules) HEMO_FAIL(ctx, HEMO_EINVAL, "out of host memory");
    }
    HemoP2Rule* rules = (HemoP2Rule*)ctx->p2rules;
    HemoP2Rule& r = rules[block];
    r.nq = nq;
    for (int q = 0; q < nq; ++q) { r.pt[q][0] = pts[2 * q]; r.pt[q][1] = pts[2 * q + 1]; r.pt[q][2] = wts[q]; }
    ctx->have_rule[block] = true;
    ctx->qrules_dirty = true;
    hemo_p2_rule_aliases(rules, ctx->have_rule, HEMO_NRULES);
    return 0;
}

static int p2_upload_constants(hemo_ctx* ctx) {
    if (!ctx->qrules_dirty) return 0;
    if (!ctx->p2rules) HEMO_FAIL(ctx, HEMO_ESTATE, "quadrature rules missing");
    for (int r = 0; r < HEMO_NRULES; ++r)
        if (!ctx->have_rule[r]) HEMO_FAIL(ctx, HEMO_ESTATE, "quadrature rule missing for a block form");
    if (!ctx->have_par) HEMO_FAIL(ctx, HEMO_ESTATE, "hemo_set_params not called");
    HEMO_CHECK_CUDA(ctx, cudaMemcpyToSymbolAsync(c_p2rules, ctx->p2rules, sizeof(HemoP2Rule) * HEMO_NRULES, 0,
                                                 cudaMemcpyHostToDevice, ctx->stream));
    HEMO_CHECK_CUDA(ctx, cudaMemcpyToSymbolAsync(c_p2frule, &ctx->frule, sizeof(HemoFacetRule), 0, cudaMemcpyHostToDevice,
                                                 ctx->stream));
    hemo_form_finalize(ctx->par);
    HEMO_CHECK_CUDA(ctx, cudaMemcpyToSymbolAsync(c_p2par, &ctx->par, sizeof(HemoForm), 0, cudaMemcpyHostToDevice, ctx->stream));
    HEMO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->qrules_dirty = false;
    return 0;
}

int hemo_p2_cell_jacobian(hemo_ctx* ctx, const double* x_dev, const double* un_dev) {
    int rc = p2_upload_constants(ctx);
    if (rc) return rc;
    const int E = ctx->E;
    const double* uh = ctx->uh ? ctx->uh : un_dev;
    int passes[4][2];
    const int np = hemo_p2_jacobian_passes((const HemoP2Rule*)ctx->p2rules, passes);
    const dim3 grid(hemo_grid(E, 128), 6);
    for (int i = 0; i < np; ++i) {
        const int r = passes[i][0];
#define P2_LAUNCH(M) k_p2_cell_jacobian<M><<<grid, 128, 0, ctx->stream>>>(E, ctx->n, r, ctx->cells, ctx->x, ctx->h, x_dev, un_dev, uh, ctx->Ae)
        switch (passes[i][1]) {
            case P2_UU: P2_LAUNCH(P2_UU); break;
            case P2_UP: P2_LAUNCH(P2_UP); break;
            case P2_PU: P2_LAUNCH(P2_PU); break;
            case P2_PP: P2_LAUNCH(P2_PP); break;
            case (P2_UP | P2_PU): P2_LAUNCH((P2_UP | P2_PU)); break;
            default: P2_LAUNCH(15); break;
        }
#undef P2_LAUNCH
        HEMO_LAUNCH_CHECK(ctx);
    }
    return 0;
}

int hemo_p2_cell_residual(hemo_ctx* ctx, const double* x_dev, const double* un_dev, const uint8_t* cellflag) {
    int rc = p2_upload_constants(ctx);
    if (rc) return rc;
    const int E = ctx->E;
    const double* uh = ctx->uh ? ctx->uh : un_dev;
    k_p2_cell_residual<<<hemo_grid(E, 128), 128, 0, ctx->stream>>>(E, ctx->n, ctx->cells, ctx->x, ctx->h, x_dev, un_dev, uh,
                                                                   cellflag, ctx->dvec, ctx->Fe);
    HEMO_LAUNCH_CHECK(ctx);
    return 0;
}

int hemo_p2_facets(hemo_ctx* ctx, int mode, const HemoFacetSet& fs, const double* x_dev, const double* un_dev,
                   const uint8_t* cellflag) {
    int rc = p2_upload_constants(ctx);
    if (rc) return rc;
    const int E = ctx->E, n = ctx->n;
    const double* uh = ctx->uh ? ctx->uh : un_dev;
    if (mode == 1)
        k_p2_facets<1><<<hemo_grid(fs.m, 64), 64, 0, ctx->stream>>>(fs.m, E, n, fs.cells, fs.mask, fs.coef, ctx->cells, ctx->x,
                                                                    ctx->h, x_dev, un_dev, uh, nullptr, nullptr, ctx->Ae);
    else
        k_p2_facets<0><<<hemo_grid(fs.m, 64), 64, 0, ctx->stream>>>(fs.m, E, n, fs.cells, fs.mask, fs.coef, ctx->cells, ctx->x,
                                                                    ctx->h, x_dev, un_dev, uh, cellflag, ctx->dvec, ctx->Fe);
    HEMO_LAUNCH_CHECK(ctx);
    return 0;
}

int hemo_p2_facet_flux(hemo_ctx* ctx, const HemoFacetSet& fs, const double* un_dev, double* partial) {
    int rc = p2_upload_constants(ctx);
    if (rc) return rc;
    k_p2_facet_flux<<<hemo_grid(fs.m, 128), 128, 0, ctx->stream>>>(fs.m, fs.cells, fs.mask, ctx->cells, ctx->x, un_dev, partial);
    HEMO_LAUNCH_CHECK(ctx);
    return 0;
}

int hemo_p2_laplace_mass(hemo_ctx* ctx) {
    k_p2_cell_laplace<<<hemo_grid(ctx->E, 256), 256, 0, ctx->stream>>>(ctx->E, ctx->cells, ctx->x, ctx->Ae, ctx->Fe);
    HEMO_LAUNCH_CHECK(ctx);
    return 0;
}
