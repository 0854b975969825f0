// Q1–Q1 quadrilateral cell and exterior-facet kernels for sm_100a: the element routines of
// q1_element.cuh wrapped in load / call / store kernels that write the same SoA element
// buffers (Ae[(a*4+b)*9 + ri*3+ci][E], Fe[a*3+comp][E]) the atomic-free gather kernels of
// assembly.cu read.
//
// Replaces the FFCx quadrilateral kernels behind assemble_matrix_block /
// assemble_vector_block (reference src/solvers/stabilized_schur.py:154,172-174) for the
// recombined transfinite mesh of src/scenarios/stenosis_pressure_structured.py:379-386.
//
// Work decomposition: at the reference quadrature (12 x 12 points for the degree-22 block
// forms) the 12 x 12 element tensor costs several 10^5 flops per cell — FP64-pipe bound by two
// orders of magnitude over its 1.2 kB of output — so the Jacobian kernel splits a cell into
// three work items (one launch each: J_uu rows {0,1}, J_uu rows {2,3}, J_up + J_pu + J_pp), each one
// thread with 32-64 accumulators in registers; every store of a warp is one coalesced 256-byte
// line.  The point set-up (geometry, tau) is recomputed per work item instead of being exchanged
// through shared memory (LDS bandwidth would cost more than the recomputation).
#include "hemo_internal.cuh"
#include "q1_element.cuh"

__constant__ HemoQuadRule c_qrules[HEMO_NRULES];
__constant__ HemoFacetRule c_qfrule;
__constant__ HemoForm c_qpar;

__device__ __forceinline__ void q1_load(Q1Cell& cd, int c, const int32_t* __restrict__ cells,
                                        const double* __restrict__ x, const double* __restrict__ h,
                                        const double* __restrict__ sol, const double* __restrict__ un,
                                        const double* __restrict__ uh, int n, int v[4]) {
    const int4 vv = reinterpret_cast<const int4*>(cells)[c];
    v[0] = vv.x; v[1] = vv.y; v[2] = vv.z; v[3] = vv.w;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const double2 xv = reinterpret_cast<const double2*>(x)[v[a]];
        cd.X[a][0] = xv.x; cd.X[a][1] = xv.y;
        const double2 uv = reinterpret_cast<const double2*>(sol)[v[a]];
        cd.U[a][0] = uv.x; cd.U[a][1] = uv.y;
        const double2 nv = reinterpret_cast<const double2*>(un)[v[a]];
        cd.N[a][0] = nv.x; cd.N[a][1] = nv.y;
        const double2 hv = reinterpret_cast<const double2*>(uh)[v[a]];
        cd.H[a][0] = hv.x; cd.H[a][1] = hv.y;
        cd.P[a] = sol[2 * (int64_t)n + v[a]];
    }
    cd.h = h[c];
    q1_prepare(cd, c_qpar);
}

template <int ITEM>
__global__ void __launch_bounds__(128)
k_q1_cell_jacobian(int E, int n, const int32_t* __restrict__ cells, const double* __restrict__ x,
                   const double* __restrict__ h, const double* __restrict__ sol,
                   const double* __restrict__ un, const double* __restrict__ uh, double* __restrict__ Ae) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= E) return;
    Q1Cell cd;
    int v[4];
    q1_load(cd, c, cells, x, h, sol, un, uh, n, v);
    const int64_t stride = E;
    double* out = Ae + c;
    auto emit = [&](int slot, double val) { out[slot * stride] = val; };
    // work item: J_uu rows {0,1} | J_uu rows {2,3} | J_up, J_pu, J_pp
    if (ITEM == 0) q1_cell_jacobian_uu<0>(cd, c_qpar, c_qrules, emit);
    else if (ITEM == 1) q1_cell_jacobian_uu<2>(cd, c_qpar, c_qrules, emit);
    else q1_cell_jacobian_p(cd, c_qpar, c_qrules, emit);
}

// Lifting of a boundary-adjacent cell, out of line and self-contained (it reloads the cell and adds
// into Fe after the residual has been stored) so that the hot path of k_q1_cell_residual keeps the
// cell data and its 12 accumulators in registers: no array of the caller has its address taken.
__device__ __noinline__ void q1_lift_device(int c, int E, int n, const int32_t* __restrict__ cells,
                                            const double* __restrict__ x, const double* __restrict__ h,
                                            const double* __restrict__ sol, const double* __restrict__ un,
                                            const double* __restrict__ uh, const double* __restrict__ dvec,
                                            double* __restrict__ Fe) {
    Q1Cell cd;
    int v[4];
    q1_load(cd, c, cells, x, h, sol, un, uh, n, v);
    double dl[4][3];
    bool any = false;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        dl[b][0] = dvec[2 * (int64_t)v[b]];
        dl[b][1] = dvec[2 * (int64_t)v[b] + 1];
        dl[b][2] = dvec[2 * (int64_t)n + v[b]];
        any = any || dl[b][0] != 0.0 || dl[b][1] != 0.0 || dl[b][2] != 0.0;
    }
    if (!any) return;
    double Fu[4][2], Fp[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) { Fu[a][0] = Fu[a][1] = 0.0; Fp[a] = 0.0; }
    q1_cell_lift(cd, c_qpar, c_qrules, dl, Fu, Fp);
    const int64_t stride = E;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        Fe[(a * 3 + 0) * stride + c] += Fu[a][0];
        Fe[(a * 3 + 1) * stride + c] += Fu[a][1];
        Fe[(a * 3 + 2) * stride + c] += Fp[a];
    }
}

__global__ void __launch_bounds__(128)
k_q1_cell_residual(int E, int n, const int32_t* __restrict__ cells, const double* __restrict__ x,
                   const double* __restrict__ h, const double* __restrict__ sol,
                   const double* __restrict__ un, const double* __restrict__ uh,
                   const uint8_t* __restrict__ cellflag, const double* __restrict__ dvec, double* __restrict__ Fe) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= E) return;
    {
        Q1Cell cd;
        int v[4];
        q1_load(cd, c, cells, x, h, sol, un, uh, n, v);
        double Fu[4][2], Fp[4];
        q1_cell_residual(cd, c_qpar, c_qrules, Fu, Fp);
        const int64_t stride = E;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            Fe[(a * 3 + 0) * stride + c] = Fu[a][0];
            Fe[(a * 3 + 1) * stride + c] = Fu[a][1];
            Fe[(a * 3 + 2) * stride + c] = Fp[a];
        }
    }
    if (cellflag != nullptr && cellflag[c]) q1_lift_device(c, E, n, cells, x, h, sol, un, uh, dvec, Fe);
}

// One thread per boundary cell of a tagged set; mode 0: residual (+ lifting) into Fe,
// mode 1: Jacobian into Ae.  Boundary-sized work: read-modify-write straight on the buffers.
template <int MODE>
__global__ void __launch_bounds__(128)
k_q1_facets(int m, int E, int n, const int32_t* __restrict__ fcells, const int32_t* __restrict__ fmask,
            hemo_facet_coef co, const int32_t* __restrict__ cells, const double* __restrict__ x,
            const double* __restrict__ h, const double* __restrict__ sol, const double* __restrict__ un,
            const double* __restrict__ uh, const uint8_t* __restrict__ cellflag, const double* __restrict__ dvec,
            double* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const int c = fcells[t];
    const int mask = fmask[t];
    Q1Cell cd;
    int v[4];
    q1_load(cd, c, cells, x, h, sol, un, uh, n, v);
    const int64_t stride = E;
    if (MODE == 1) {
        q1_cell_facets(cd, c_qpar, c_qfrule, co, mask, false, true,
                       [&](int, int, double) {},
                       [&](int a, int b, int ri, int ci, double val) {
                           out[((a * 4 + b) * 9 + ri * 3 + ci) * stride + c] += val;
                       });
    } else {
        double dl[4][3];
        bool lift = false;
        if (cellflag != nullptr && cellflag[c]) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                dl[b][0] = dvec[2 * (int64_t)v[b]];
                dl[b][1] = dvec[2 * (int64_t)v[b] + 1];
                dl[b][2] = dvec[2 * (int64_t)n + v[b]];
                lift = lift || dl[b][0] != 0.0 || dl[b][1] != 0.0 || dl[b][2] != 0.0;
            }
        }
        double Fu[4][2];
#pragma unroll
        for (int a = 0; a < 4; ++a) Fu[a][0] = Fu[a][1] = 0.0;
        q1_cell_facets(cd, c_qpar, c_qfrule, co, mask, true, lift,
                       [&](int a, int k, double val) { Fu[a][k] += val; },
                       [&](int a, int b, int ri, int ci, double val) { Fu[a][ri] += val * dl[b][ci]; });
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            out[(a * 3 + 0) * stride + c] += Fu[a][0];
            out[(a * 3 + 1) * stride + c] += Fu[a][1];
        }
    }
}

__global__ void k_q1_facet_flux(int m, const int32_t* __restrict__ fcells, const int32_t* __restrict__ fmask,
                                const int32_t* __restrict__ cells, const double* __restrict__ x,
                                const double* __restrict__ un, double* __restrict__ partial) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const int c = fcells[t];
    Q1Cell cd;
    for (int a = 0; a < 4; ++a) {
        const int v = cells[4 * (int64_t)c + a];
        cd.X[a][0] = x[2 * (int64_t)v]; cd.X[a][1] = x[2 * (int64_t)v + 1];
        cd.N[a][0] = un[2 * (int64_t)v]; cd.N[a][1] = un[2 * (int64_t)v + 1];
    }
    partial[t] = q1_cell_flux(cd, fmask[t]);
}

__global__ void k_q1_cell_laplace(int E, const int32_t* __restrict__ cells, const double* __restrict__ x,
                                  double* __restrict__ Ke /*[16][E]*/, double* __restrict__ Me /*[4][E]*/) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= E) return;
    Q1Cell cd;
    for (int a = 0; a < 4; ++a) {
        const int v = cells[4 * (int64_t)c + a];
        cd.X[a][0] = x[2 * (int64_t)v]; cd.X[a][1] = x[2 * (int64_t)v + 1];
    }
    double K[4][4], M[4];
    q1_cell_laplace_mass(cd, K, M);
    for (int a = 0; a < 4; ++a) {
        for (int b = 0; b < 4; ++b) Ke[(int64_t)(a * 4 + b) * E + c] = K[a][b];
        Me[(int64_t)a * E + c] = M[a];
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int q1_upload_constants(hemo_ctx* ctx) {
    if (!ctx->qrules_dirty) return 0;
    if (!ctx->qrules) HEMO_FAIL(ctx, HEMO_ESTATE, "quadrature rules missing");
    for (int r = 0; r < HEMO_NRULES; ++r)
        if (!ctx->have_rule[r]) HEMO_FAIL(ctx, HEMO_ESTATE, "quadrature rule missing for a block form");
    if (!ctx->have_par) HEMO_FAIL(ctx, HEMO_ESTATE, "hemo_set_params not called");
    HEMO_CHECK_CUDA(ctx, cudaMemcpyToSymbolAsync(c_qrules, ctx->qrules, sizeof(HemoQuadRule) * HEMO_NRULES, 0,
                                                 cudaMemcpyHostToDevice, ctx->stream));
    HEMO_CHECK_CUDA(ctx, cudaMemcpyToSymbolAsync(c_qfrule, &ctx->frule, sizeof(HemoFacetRule), 0,
                                                 cudaMemcpyHostToDevice, ctx->stream));
    hemo_form_finalize(ctx->par);
    HEMO_CHECK_CUDA(ctx, cudaMemcpyToSymbolAsync(c_qpar, &ctx->par, sizeof(HemoForm), 0,
                                                 cudaMemcpyHostToDevice, ctx->stream));
    // the host copies above are read when the (pageable-memory) copy is staged, i.e. before return
    ctx->qrules_dirty = false;
    return 0;
}

int hemo_q1_cell_jacobian(hemo_ctx* ctx, const double* x_dev, const double* un_dev) {
    int rc = q1_upload_constants(ctx);
    if (rc) return rc;
    const int E = ctx->E;
    const double* uh = ctx->uh ? ctx->uh : un_dev;
    const int grid = hemo_grid(E, 128);
    k_q1_cell_jacobian<0><<<grid, 128, 0, ctx->stream>>>(E, ctx->n, ctx->cells, ctx->x, ctx->h, x_dev, un_dev, uh, ctx->Ae);
    HEMO_LAUNCH_CHECK(ctx);
    k_q1_cell_jacobian<1><<<grid, 128, 0, ctx->stream>>>(E, ctx->n, ctx->cells, ctx->x, ctx->h, x_dev, un_dev, uh, ctx->Ae);
    HEMO_LAUNCH_CHECK(ctx);
    k_q1_cell_jacobian<2><<<grid, 128, 0, ctx->stream>>>(E, ctx->n, ctx->cells, ctx->x, ctx->h, x_dev, un_dev, uh, ctx->Ae);
    HEMO_LAUNCH_CHECK(ctx);
    return 0;
}

int hemo_q1_cell_residual(hemo_ctx* ctx, const double* x_dev, const double* un_dev, const uint8_t* cellflag) {
    int rc = q1_upload_constants(ctx);
    if (rc) return rc;
    const int E = ctx->E;
    const double* uh = ctx->uh ? ctx->uh : un_dev;
    // (a 3-CTA/SM register budget, 168 registers, was measured at the same 5.9 ms per million cells:
    // the kernel is bound by the FP64 pipe, not by latency)
    k_q1_cell_residual<<<hemo_grid(E, 128), 128, 0, ctx->stream>>>(E, ctx->n, ctx->cells, ctx->x, ctx->h, x_dev, un_dev,
                                                                   uh, cellflag, ctx->dvec, ctx->Fe);
    HEMO_LAUNCH_CHECK(ctx);
    return 0;
}

int hemo_q1_facets(hemo_ctx* ctx, int mode, const HemoFacetSet& fs, const double* x_dev, const double* un_dev,
                   const uint8_t* cellflag) {
    int rc = q1_upload_constants(ctx);
    if (rc) return rc;
    const int E = ctx->E, n = ctx->n;
    const double* uh = ctx->uh ? ctx->uh : un_dev;
    if (mode == 1)
        k_q1_facets<1><<<hemo_grid(fs.m, 128), 128, 0, ctx->stream>>>(fs.m, E, n, fs.cells, fs.mask, fs.coef, ctx->cells,
                                                                      ctx->x, ctx->h, x_dev, un_dev, uh, nullptr, nullptr,
                                                                      ctx->Ae);
    else
        k_q1_facets<0><<<hemo_grid(fs.m, 128), 128, 0, ctx->stream>>>(fs.m, E, n, fs.cells, fs.mask, fs.coef, ctx->cells,
                                                                      ctx->x, ctx->h, x_dev, un_dev, uh, cellflag, ctx->dvec,
                                                                      ctx->Fe);
    HEMO_LAUNCH_CHECK(ctx);
    return 0;
}

int hemo_q1_facet_flux(hemo_ctx* ctx, const HemoFacetSet& fs, const double* un_dev, double* partial) {
    k_q1_facet_flux<<<hemo_grid(fs.m, 128), 128, 0, ctx->stream>>>(fs.m, fs.cells, fs.mask, ctx->cells, ctx->x, un_dev,
                                                                   partial);
    HEMO_LAUNCH_CHECK(ctx);
    return 0;
}

int hemo_q1_laplace_mass(hemo_ctx* ctx) {
    k_q1_cell_laplace<<<hemo_grid(ctx->E, 256), 256, 0, ctx->stream>>>(ctx->E, ctx->cells, ctx->x, ctx->Ae, ctx->Fe);
    HEMO_LAUNCH_CHECK(ctx);
    return 0;
}
