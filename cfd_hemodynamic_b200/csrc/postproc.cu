// Per-step post-processing on the device: what the reference computes
// on the host right after `solveStep` in `Scenario.solve` — so that a device-resident time loop
// leaves no per-step host work and no per-step D2H copy of the fields:
//   * wall shear stress vector  `solver.assemble_wss()`  src/solverBase.py:144-195, src/scenario.py:258
//   * early-stop norms          src/scenario.py:268-304
//   * boundary force (drag/lift integrals of the DFG benchmark)  src/scenarios/dfg_1.py:183-211
//   * L2 norms of the final fields  src/scenario.py:315-324
// P1 triangles and Q1 quadrilaterals; reductions are fixed-order (bitwise reproducible).
#include "hemo_internal.cuh"
#include "q1_element.cuh"
#include "tet_items.cuh"

#define PP_THREADS 256
#define PP_BLOCKS 592     // 4 per SM on 148 SMs

__device__ __forceinline__ double pp_block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;
}

__device__ __forceinline__ double pp_block_max(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_down_sync(0xffffffffu, r, o));
    }
    __syncthreads();
    return r;
}

// ---- geometry of one boundary facet -------------------------------------------------
// Tangential traction Tt = T - (T.n) n, T = -2 mu eps(u) n, at facet parameter s, and the two
// facet vertices (local indices).  NV = 3: gradients constant per cell; NV = 4: evaluated at s.
template <int NV>
struct FacetCell {
    double X[NV][2], U[NV][2], P[NV];
    int v[NV];
};

template <int NV>
__device__ __forceinline__ void pp_load(FacetCell<NV>& fc, int c, const int32_t* __restrict__ cells,
                                        const double* __restrict__ x, const double* __restrict__ sol, int n) {
#pragma unroll
    for (int a = 0; a < NV; ++a) {
        const int v = cells[NV * (int64_t)c + a];
        fc.v[a] = v;
        fc.X[a][0] = x[2 * (int64_t)v]; fc.X[a][1] = x[2 * (int64_t)v + 1];
        fc.U[a][0] = sol[2 * (int64_t)v]; fc.U[a][1] = sol[2 * (int64_t)v + 1];
        fc.P[a] = sol[2 * (int64_t)n + v];
    }
}

// outward unit normal, length and end vertices of local facet lf
template <int NV>
__device__ __forceinline__ void pp_facet(const FacetCell<NV>& fc, int lf, int& va, int& vb, double nr[2], double& len) {
    double ix, iy;     // a point inside the cell: opposite vertex (triangle) / centroid (quadrilateral)
    if constexpr (NV == 3) {
        va = (lf == 0) ? 1 : 0;
        vb = (lf == 2) ? 1 : 2;
        ix = fc.X[lf][0]; iy = fc.X[lf][1];
    } else {
        q1_facet_verts(lf, va, vb);
        ix = 0.25 * (fc.X[0][0] + fc.X[1][0] + fc.X[2][0] + fc.X[3][0]);
        iy = 0.25 * (fc.X[0][1] + fc.X[1][1] + fc.X[2][1] + fc.X[3][1]);
    }
    const double tx = fc.X[vb][0] - fc.X[va][0], ty = fc.X[vb][1] - fc.X[va][1];
    len = sqrt(tx * tx + ty * ty);
    double nx = ty / len, ny = -tx / len;
    const double side = nx * (0.5 * (fc.X[va][0] + fc.X[vb][0]) - ix) + ny * (0.5 * (fc.X[va][1] + fc.X[vb][1]) - iy);
    if (side < 0.0) { nx = -nx; ny = -ny; }
    nr[0] = nx; nr[1] = ny;
}

// G[i][j] = d_i u_j at parameter s of local facet lf
template <int NV>
__device__ __forceinline__ void pp_grad(const FacetCell<NV>& fc, int lf, double s, double G[2][2]) {
    double g[NV][2];
    if constexpr (NV == 3) {
        const double j00 = fc.X[1][0] - fc.X[0][0], j01 = fc.X[2][0] - fc.X[0][0];
        const double j10 = fc.X[1][1] - fc.X[0][1], j11 = fc.X[2][1] - fc.X[0][1];
        const double det = j00 * j11 - j01 * j10;
        const double i00 = j11 / det, i01 = -j01 / det, i10 = -j10 / det, i11 = j00 / det;
        g[1][0] = i00; g[1][1] = i01; g[2][0] = i10; g[2][1] = i11;
        g[0][0] = -(i00 + i10); g[0][1] = -(i01 + i11);
    } else {
        Q1Cell qc;
#pragma unroll
        for (int a = 0; a < 4; ++a) { qc.X[a][0] = fc.X[a][0]; qc.X[a][1] = fc.X[a][1]; }
        double xi, eta;
        q1_facet_ref(lf, s, xi, eta);
        Q1Geom ge;
        q1_geom(qc, xi, eta, ge);
#pragma unroll
        for (int a = 0; a < 4; ++a) { g[a][0] = ge.g[a][0]; g[a][1] = ge.g[a][1]; }
    }
    G[0][0] = G[0][1] = G[1][0] = G[1][1] = 0.0;
#pragma unroll
    for (int a = 0; a < NV; ++a) {
        G[0][0] += g[a][0] * fc.U[a][0]; G[0][1] += g[a][0] * fc.U[a][1];
        G[1][0] += g[a][1] * fc.U[a][0]; G[1][1] += g[a][1] * fc.U[a][1];
    }
}

// 3-point Gauss rule on [0,1] (exact for the P1 terms, accurate for the rational Q1 gradients)
__constant__ double c_pp_s[3] = {0.5 - 0.5 * 0.7745966692414834, 0.5, 0.5 + 0.5 * 0.7745966692414834};
__constant__ double c_pp_w[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};

// shear_stress_i += (1/|F|) int_F phi_i (T - (T.n) n) ds, T = -2 mu eps(u) n  (solverBase.py:144-195).
// Every boundary vertex of a 2-D mesh receives exactly two contributions, and a + b is
// commutative in floating point, so the atomic adds are order-independent.
template <int NV>
__global__ void k_wss(int m, int n, const int32_t* __restrict__ fcells, const int32_t* __restrict__ fmask,
                      const int32_t* __restrict__ cells, const double* __restrict__ x,
                      const double* __restrict__ sol, double mu, double* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    FacetCell<NV> fc;
    pp_load<NV>(fc, fcells[t], cells, x, sol, n);
    const int mask = fmask[t];
    for (int lf = 0; lf < NV; ++lf) {
        if (!(mask & (1 << lf))) continue;
        int va, vb;
        double nr[2], len;
        pp_facet<NV>(fc, lf, va, vb, nr, len);
        double ta[2] = {0, 0}, tb[2] = {0, 0};
        for (int q = 0; q < 3; ++q) {
            const double s = c_pp_s[q], w = c_pp_w[q];
            double G[2][2];
            pp_grad<NV>(fc, lf, s, G);
            const double e01 = 0.5 * (G[0][1] + G[1][0]);
            const double T0 = -2.0 * mu * (G[0][0] * nr[0] + e01 * nr[1]);
            const double T1 = -2.0 * mu * (e01 * nr[0] + G[1][1] * nr[1]);
            const double tn = T0 * nr[0] + T1 * nr[1];
            const double t0 = T0 - tn * nr[0], t1 = T1 - tn * nr[1];
            ta[0] += w * (1.0 - s) * t0; ta[1] += w * (1.0 - s) * t1;
            tb[0] += w * s * t0; tb[1] += w * s * t1;
        }
        atomicAdd(&out[2 * (int64_t)fc.v[va]], ta[0]);
        atomicAdd(&out[2 * (int64_t)fc.v[va] + 1], ta[1]);
        atomicAdd(&out[2 * (int64_t)fc.v[vb]], tb[0]);
        atomicAdd(&out[2 * (int64_t)fc.v[vb] + 1], tb[1]);
    }
}

// Drag / lift integrals of dfg_1.py:183-202 with n = -FacetNormal, t = (n_y, -n_x):
//   F_D = int mu d(u.t)/dn n_y - p n_x ds,   F_L = -int mu d(u.t)/dn n_x + p n_y ds
template <int NV>
__global__ void k_boundary_force(int m, int n, const int32_t* __restrict__ fcells, const int32_t* __restrict__ fmask,
                                 const int32_t* __restrict__ cells, const double* __restrict__ x,
                                 const double* __restrict__ sol, double mu, double* __restrict__ partial /*[2][m]*/) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    FacetCell<NV> fc;
    pp_load<NV>(fc, fcells[t], cells, x, sol, n);
    const int mask = fmask[t];
    double fd = 0.0, fl = 0.0;
    for (int lf = 0; lf < NV; ++lf) {
        if (!(mask & (1 << lf))) continue;
        int va, vb;
        double nout[2], len;
        pp_facet<NV>(fc, lf, va, vb, nout, len);
        const double nx = -nout[0], ny = -nout[1];
        const double tx = ny, ty = -nx;
        for (int q = 0; q < 3; ++q) {
            const double s = c_pp_s[q], w = c_pp_w[q] * len;
            double G[2][2];
            pp_grad<NV>(fc, lf, s, G);
            // grad(u_t) . n = n_i G_ij t_j
            const double dut = nx * (G[0][0] * tx + G[0][1] * ty) + ny * (G[1][0] * tx + G[1][1] * ty);
            const double p = (1.0 - s) * fc.P[va] + s * fc.P[vb];
            fd += w * (mu * dut * ny - p * nx);
            fl -= w * (mu * dut * nx + p * ny);
        }
    }
    partial[t] = fd;
    partial[m + t] = fl;
}

__global__ void k_pp_sum2(int m, const double* __restrict__ partial, double* __restrict__ out) {
    __shared__ double sh[32];
    const double* p = partial + (int64_t)blockIdx.x * m;
    double acc = 0.0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) acc += p[i];
    acc = pp_block_sum(acc, sh);
    if (threadIdx.x == 0) out[blockIdx.x] = acc;
}

// partial[b] = max |u - un| over the block's slice, partial[gridDim.x + b] = max |u|
__global__ void __launch_bounds__(PP_THREADS)
k_inf_norms(int64_t N, const double* __restrict__ u, const double* __restrict__ un, double* __restrict__ partial) {
    __shared__ double sh[32];
    double d = 0.0, a = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = u[i];
        d = fmax(d, fabs(v - un[i]));
        a = fmax(a, fabs(v));
    }
    d = pp_block_max(d, sh);
    a = pp_block_max(a, sh);
    if (threadIdx.x == 0) { partial[blockIdx.x] = d; partial[gridDim.x + blockIdx.x] = a; }
}

__global__ void k_pp_max2(int nblk, const double* __restrict__ partial, double* __restrict__ out) {
    __shared__ double sh[32];
    const double* p = partial + (int64_t)blockIdx.x * nblk;
    double acc = 0.0;
    for (int i = threadIdx.x; i < nblk; i += blockDim.x) acc = fmax(acc, p[i]);
    acc = pp_block_max(acc, sh);
    if (threadIdx.x == 0) out[blockIdx.x] = acc;
}

// int |f|^2 dx per cell, f a bs-component nodal field (inner(u,u)*dx, scenario.py:315-324).
// Triangles: exact P1 mass matrix; quadrilaterals: 3 x 3 Gauss (exact on affine cells).
template <int NV>
__global__ void __launch_bounds__(PP_THREADS)
k_l2_partial(int E, int bs, const int32_t* __restrict__ cells, const double* __restrict__ x,
             const double* __restrict__ f, double* __restrict__ partial) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < E; c += gridDim.x * blockDim.x) {
        int v[NV];
        double X[NV][2];
#pragma unroll
        for (int a = 0; a < NV; ++a) {
            v[a] = cells[NV * (int64_t)c + a];
            X[a][0] = x[2 * (int64_t)v[a]]; X[a][1] = x[2 * (int64_t)v[a] + 1];
        }
        for (int k = 0; k < bs; ++k) {
            double F[NV];
#pragma unroll
            for (int a = 0; a < NV; ++a) F[a] = f[(int64_t)v[a] * bs + k];
            if constexpr (NV == 3) {
                const double det = fabs((X[1][0] - X[0][0]) * (X[2][1] - X[0][1]) - (X[2][0] - X[0][0]) * (X[1][1] - X[0][1]));
                const double s1 = F[0] + F[1] + F[2];
                acc += det / 24.0 * (s1 * s1 + F[0] * F[0] + F[1] * F[1] + F[2] * F[2]);
            } else {
                Q1Cell qc;
#pragma unroll
                for (int a = 0; a < 4; ++a) { qc.X[a][0] = X[a][0]; qc.X[a][1] = X[a][1]; }
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) {
                        Q1Geom ge;
                        q1_geom(qc, c_pp_s[i], c_pp_s[j], ge);
                        double fv = 0.0;
#pragma unroll
                        for (int a = 0; a < 4; ++a) fv += ge.phi[a] * F[a];
                        acc += c_pp_w[i] * c_pp_w[j] * ge.adet * fv * fv;
                    }
            }
        }
    }
    acc = pp_block_sum(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// ---- tetrahedra (bodies in tet_items.cuh) ------------------------------------------------------
// A boundary vertex of a tetrahedral mesh collects the traction of every tagged facet around it.  Like every other
// reduction of the library this one is atomic-free with a fixed order: pass 1 stores the contributions of each tagged
// cell to its four vertices, pass 2 sums them per node in the order of the node's cell list (vseg) — bitwise
// reproducible.
__global__ void k_tet_wss_items(int m, const int32_t* __restrict__ fcells, const int32_t* __restrict__ fmask,
                                const int32_t* __restrict__ cells, const double* __restrict__ x,
                                const double* __restrict__ sol, double mu, double* __restrict__ tmp /* [m][4][3] */) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    double acc[4][3];
    for (int a = 0; a < 4; ++a) acc[a][0] = acc[a][1] = acc[a][2] = 0.0;
    int v[4];
    const int c = fcells[t];
    for (int a = 0; a < 4; ++a) v[a] = cells[4 * (int64_t)c + a];
    tet_wss_item(t, fcells, fmask, cells, x, sol, mu, [&](int node, int k, double val) {
        for (int a = 0; a < 4; ++a)
            if (v[a] == node) { acc[a][k] += val; break; }
    });
    for (int a = 0; a < 4; ++a)
        for (int k = 0; k < 3; ++k) tmp[(int64_t)t * 12 + a * 3 + k] = acc[a][k];
}

__global__ void k_set_cell_index(int m, const int32_t* __restrict__ fcells, int32_t* __restrict__ cell2t) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < m) cell2t[fcells[t]] = t;
}

__global__ void k_tet_wss_gather(int n, const int32_t* __restrict__ vseg_ptr, const int32_t* __restrict__ vseg_src,
                                 const int32_t* __restrict__ cell2t, const double* __restrict__ tmp, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int q = vseg_ptr[i]; q < vseg_ptr[i + 1]; ++q) {
        const int src = vseg_src[q];
        const int t = cell2t[src >> 2];
        if (t < 0) continue;
        const double* p = tmp + (int64_t)t * 12 + (src & 3) * 3;
        a0 += p[0]; a1 += p[1]; a2 += p[2];
    }
    out[3 * (int64_t)i] = a0; out[3 * (int64_t)i + 1] = a1; out[3 * (int64_t)i + 2] = a2;
}

__global__ void __launch_bounds__(PP_THREADS)
k_tet_l2_partial(int E, int bs, const int32_t* __restrict__ cells, const double* __restrict__ x,
                 const double* __restrict__ f, double* __restrict__ partial) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < E; c += gridDim.x * blockDim.x)
        acc += tet_l2_item(c, bs, cells, x, f);
    acc = pp_block_sum(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int pp_fetch(hemo_ctx* ctx, int count, double* out_host) {
    HEMO_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->red_host, ctx->red_out, sizeof(double) * count, cudaMemcpyDeviceToHost,
                                         ctx->stream));
    HEMO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < count; ++i) out_host[i] = ctx->red_host[i];
    return 0;
}

extern "C" int hemo_wall_shear_stress(hemo_ctx* ctx, int set_id, const double* x_dev, double* wss_dev) {
    if (!ctx || set_id < 0 || set_id >= HEMO_MAX_FACET_SETS || !x_dev || !wss_dev) return HEMO_EINVAL;
    if (!ctx->cells || !ctx->have_par) HEMO_FAIL(ctx, HEMO_ESTATE, "mesh / params not set");
    const HemoFacetSet& fs = ctx->fsets[set_id];
    HEMO_CHECK_CUDA(ctx, cudaMemsetAsync(wss_dev, 0, sizeof(double) * ctx->dim * ctx->n, ctx->stream));
    if (fs.m == 0) return 0;
    const int grid = hemo_grid(fs.m, 128);
    if (ctx->dim == 3) {
        if (!ctx->vseg_ptr) HEMO_FAIL(ctx, HEMO_ESTATE, "node graph not set");
        int rc;
        if (ctx->wss_set != set_id || ctx->wss_version != ctx->fset_version) {
            if ((rc = hemo_alloc(ctx, &ctx->wss_cell2t, (size_t)ctx->E))) return rc;
            if ((rc = hemo_alloc(ctx, &ctx->wss_tmp, (size_t)fs.m * 12))) return rc;
            HEMO_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->wss_cell2t, 0xff, sizeof(int32_t) * ctx->E, ctx->stream));
            k_set_cell_index<<<grid, 128, 0, ctx->stream>>>(fs.m, fs.cells, ctx->wss_cell2t);
            HEMO_LAUNCH_CHECK(ctx);
            ctx->wss_set = set_id;
            ctx->wss_version = ctx->fset_version;
        }
        k_tet_wss_items<<<grid, 128, 0, ctx->stream>>>(fs.m, fs.cells, fs.mask, ctx->cells, ctx->x, x_dev, ctx->par.mu, ctx->wss_tmp);
        HEMO_LAUNCH_CHECK(ctx);
        k_tet_wss_gather<<<hemo_grid(ctx->n, 256), 256, 0, ctx->stream>>>(ctx->n, ctx->vseg_ptr, ctx->vseg_src, ctx->wss_cell2t,
                                                                         ctx->wss_tmp, wss_dev);
    } else if (ctx->nv == 4)
        k_wss<4><<<grid, 128, 0, ctx->stream>>>(fs.m, ctx->n, fs.cells, fs.mask, ctx->cells, ctx->x, x_dev, ctx->par.mu, wss_dev);
    else
        k_wss<3><<<grid, 128, 0, ctx->stream>>>(fs.m, ctx->n, fs.cells, fs.mask, ctx->cells, ctx->x, x_dev, ctx->par.mu, wss_dev);
    HEMO_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int hemo_boundary_force(hemo_ctx* ctx, int set_id, const double* x_dev, double* force_host) {
    if (!ctx || set_id < 0 || set_id >= HEMO_MAX_FACET_SETS || !x_dev || !force_host) return HEMO_EINVAL;
    HEMO_2D_ONLY(ctx, "boundary force");
    if (!ctx->cells || !ctx->have_par) HEMO_FAIL(ctx, HEMO_ESTATE, "mesh / params not set");
    const HemoFacetSet& fs = ctx->fsets[set_id];
    force_host[0] = force_host[1] = 0.0;
    if (fs.m == 0) return 0;
    int rc = hemo_ensure_reduce(ctx, (size_t)2 * fs.m, 8);
    if (rc) return rc;
    const int grid = hemo_grid(fs.m, 128);
    if (ctx->nv == 4)
        k_boundary_force<4><<<grid, 128, 0, ctx->stream>>>(fs.m, ctx->n, fs.cells, fs.mask, ctx->cells, ctx->x, x_dev,
                                                           ctx->par.mu, ctx->red_partial);
    else
        k_boundary_force<3><<<grid, 128, 0, ctx->stream>>>(fs.m, ctx->n, fs.cells, fs.mask, ctx->cells, ctx->x, x_dev,
                                                           ctx->par.mu, ctx->red_partial);
    HEMO_LAUNCH_CHECK(ctx);
    k_pp_sum2<<<2, 256, 0, ctx->stream>>>(fs.m, ctx->red_partial, ctx->red_out);
    HEMO_LAUNCH_CHECK(ctx);
    return pp_fetch(ctx, 2, force_host);
}

extern "C" int hemo_early_stop_norms(hemo_ctx* ctx, int64_t n, const double* u_dev, const double* un_dev,
                                     double* norms_host) {
    if (!ctx || !u_dev || !un_dev || !norms_host || n <= 0) return HEMO_EINVAL;
    int rc = hemo_ensure_reduce(ctx, (size_t)2 * PP_BLOCKS, 8);
    if (rc) return rc;
    int64_t g = (n + PP_THREADS - 1) / PP_THREADS;
    if (g > PP_BLOCKS) g = PP_BLOCKS;
    k_inf_norms<<<(int)g, PP_THREADS, 0, ctx->stream>>>(n, u_dev, un_dev, ctx->red_partial);
    HEMO_LAUNCH_CHECK(ctx);
    k_pp_max2<<<2, 256, 0, ctx->stream>>>((int)g, ctx->red_partial, ctx->red_out);
    HEMO_LAUNCH_CHECK(ctx);
    return pp_fetch(ctx, 2, norms_host);
}

extern "C" int hemo_l2_norm_sq(hemo_ctx* ctx, int bs, const double* f_dev, double* out_host) {
    if (!ctx || !f_dev || !out_host || (bs != 1 && bs != ctx->dim)) return HEMO_EINVAL;
    if (!ctx->cells) HEMO_FAIL(ctx, HEMO_ESTATE, "mesh not set");
    int rc = hemo_ensure_reduce(ctx, (size_t)PP_BLOCKS, 8);
    if (rc) return rc;
    int g = hemo_grid(ctx->E, PP_THREADS);
    if (g > PP_BLOCKS) g = PP_BLOCKS;
    if (ctx->dim == 3)
        k_tet_l2_partial<<<g, PP_THREADS, 0, ctx->stream>>>(ctx->E, bs, ctx->cells, ctx->x, f_dev, ctx->red_partial);
    else if (ctx->nv == 4)
        k_l2_partial<4><<<g, PP_THREADS, 0, ctx->stream>>>(ctx->E, bs, ctx->cells, ctx->x, f_dev, ctx->red_partial);
    else
        k_l2_partial<3><<<g, PP_THREADS, 0, ctx->stream>>>(ctx->E, bs, ctx->cells, ctx->x, f_dev, ctx->red_partial);
    HEMO_LAUNCH_CHECK(ctx);
    k_pp_sum2<<<1, 256, 0, ctx->stream>>>(g, ctx->red_partial, ctx->red_out);
    HEMO_LAUNCH_CHECK(ctx);
    return pp_fetch(ctx, 1, out_host);
}
