// Curl-curl (rotational) formulation on P1-P1 simplices: the cell and exterior-facet integrals of
// src/solvers/stabilized_schur_pressurebc.py:85-160 (cells) and :189-201 (weak pressure + curl-form Nitsche), selected
// with hemo_set_formulation(ctx, HEMO_FORM_CURLCURL).  Triangles and tetrahedra share the code (template <int D>,
// csrc/curlcurl_element.cuh); everything after the element tensors — gathers into the reference's CSR, Dirichlet
// rows / lifting / set_bc, SpMV, the preconditioner and FGMRES — is the standard path's.
//
// FFCx-style: the full integrand at every point of each block form's own rule (rules shared by several block
// forms are integrated once, alias table); one thread per (cell, test node) so that the (D+1) x (D+1)^2 rows of the
// element Jacobian it owns stay in registers.  Not moment-factorised or tuned: a first wiring, FP64-pipe bound.
#include "hemo_internal.cuh"
#include "curlcurl_element.cuh"

enum { CC_JAC_ALL = 0, CC_JAC_FLAGGED = 1, CC_JAC_NONE = 2 };   // = the TET_JAC_* modes of assembly_tet.cu

template <int D>
struct CcRules {
    SimplexRule<D> r[HEMO_NRULES];
    int alias[HEMO_NRULES];          // lowest block id with an identical rule
};

struct hemo_cc_state {
    void* dev = nullptr;             // CcRules<D> on the device
    int dim = 0;
    int64_t version = -1;            // ctx->rule_version the tables were built from
};

// rule tables of the tetrahedron path (assembly_tet.cu)
int hemo_tet_get_rules(hemo_ctx* ctx, const SimplexRule<3>** rules, const int** alias, const SimplexFacetRule<3>** frule);

template <int D>
__device__ __forceinline__ void cc_load_cell(SimplexCell<D>& cd, int c, int n, const int32_t* __restrict__ cells,
                                             const double* __restrict__ x, const double* __restrict__ h,
                                             const double* __restrict__ sol, const double* __restrict__ un,
                                             const double* fb, int v[D + 1]) {
    constexpr int NV = D + 1;
    double X[NV][D];
#pragma unroll
    for (int a = 0; a < NV; ++a) {
        v[a] = cells[NV * (int64_t)c + a];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            X[a][k] = x[D * (int64_t)v[a] + k];
            cd.U[a][k] = sol[D * (int64_t)v[a] + k];
            cd.N[a][k] = un[D * (int64_t)v[a] + k];
        }
        cd.P[a] = sol[D * (int64_t)n + v[a]];
    }
#pragma unroll
    for (int k = 0; k < D; ++k) cd.fbody[k] = fb[k];
    cd.h = h[c];
    simplex_geometry<D>(cd, X);
}

// blockIdx.y = test node a: F_u[a][:], F_p[a] and — where jac_mode asks for it — the rows of node a of the element
// Jacobian, SoA layouts of the standard path: Fe[a*(D+1) + comp][E], Ae[(a*NV + b)*(D+1)^2 + ri*(D+1) + ci][E].
template <int D>
__global__ void __launch_bounds__(128)
k_cc_cells(int E, int n, const int32_t* __restrict__ cells, const double* __restrict__ x, const double* __restrict__ h,
           const double* __restrict__ sol, const double* __restrict__ un, HemoForm par, double f0, double f1, double f2,
           const CcRules<D>* __restrict__ rules, int jac_mode, const uint8_t* __restrict__ jac_cells,
           double* __restrict__ Ae, double* __restrict__ Fe) {
    constexpr int NV = D + 1, NL = NV * NV, C = D + 1;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = blockIdx.y;
    if (c >= E) return;
    const double fb[3] = {f0, f1, f2};
    SimplexCell<D> cd;
    int v[NV];
    cc_load_cell<D>(cd, c, n, cells, x, h, sol, un, fb, v);
    const bool want_jac = jac_mode == CC_JAC_ALL || (jac_mode == CC_JAC_FLAGGED && jac_cells[c]);
    double Fu[NV][D], Fp[NV], J[C * NL];
#pragma unroll
    for (int b = 0; b < NV; ++b) {
        Fp[b] = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) Fu[b][k] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < C * NL; ++i) J[i] = 0.0;
    const int bit[HEMO_NRULES] = {CC_FU, CC_FP, CC_UU, CC_UP, CC_PU, CC_PP};     // block ids HEMO_Q_FU .. HEMO_Q_PP
    for (int b = 0; b < HEMO_NRULES; ++b) {
        if (rules->alias[b] != b) continue;
        int mask = 0;
        for (int bb = b; bb < HEMO_NRULES; ++bb)
            if (rules->alias[bb] == b && (bb < 2 || want_jac)) mask |= bit[bb];
        if (mask) curlcurl_cell_ex<D>(cd, par, rules->r[b], mask, a, Fu, Fp, J);
    }
    const int64_t stride = E;
#pragma unroll
    for (int k = 0; k < D; ++k) Fe[(int64_t)(a * C + k) * stride + c] = Fu[a][k];
    Fe[(int64_t)(a * C + D) * stride + c] = Fp[a];
    if (!want_jac) return;
    for (int b = 0; b < NV; ++b)
#pragma unroll
        for (int ri = 0; ri < C; ++ri)
#pragma unroll
            for (int ci = 0; ci < C; ++ci) {
                const int col = ci < D ? b * D + ci : D * NV + b;
                Ae[(int64_t)((a * NV + b) * C * C + ri * C + ci) * stride + c] = J[ri * NL + col];
            }
}

// one thread per boundary cell of a tagged set: adds the facet residual into Fe and its derivative into Ae
template <int D>
__global__ void __launch_bounds__(128)
k_cc_facets(int m, int E, int n, const int32_t* __restrict__ fcells, const int32_t* __restrict__ fmask, hemo_facet_coef co,
            SimplexFacetRule<D> fr, const int32_t* __restrict__ cells, const double* __restrict__ x,
            const double* __restrict__ h, const double* __restrict__ sol, const double* __restrict__ un, HemoForm par,
            int jac_mode, const uint8_t* __restrict__ jac_cells, double* __restrict__ Ae, double* __restrict__ Fe) {
    constexpr int NV = D + 1, C = D + 1;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const int c = fcells[t];
    const int mask = fmask[t];
    const double fb[3] = {0.0, 0.0, 0.0};
    SimplexCell<D> cd;
    int v[NV];
    cc_load_cell<D>(cd, c, n, cells, x, h, sol, un, fb, v);
    const bool want_jac = jac_mode == CC_JAC_ALL || (jac_mode == CC_JAC_FLAGGED && jac_cells[c]);
    const int64_t stride = E;
    auto res = [&](int a, int k, double val) { Fe[(int64_t)(a * C + k) * stride + c] += val; };
    auto jac = [&](int a, int b, int k, int l, double val) {
        Ae[(int64_t)((a * NV + b) * C * C + k * C + l) * stride + c] += val;
    };
    for (int lf = 0; lf < NV; ++lf) {
        if (!(mask & (1 << lf))) continue;
        if (want_jac) curlcurl_facet<D, true, true>(cd, par, co, fr, lf, res, jac);
        else curlcurl_facet<D, true, false>(cd, par, co, fr, lf, res, jac);
    }
}

// lifting of assemble_vector_block(..., x0 = x, alpha = -1) on Dirichlet-adjacent cells: Fe += Ae d, d = g - x on
// constrained dofs (dvec in the global layout [u interleaved (D n) | p (n)])
template <int D>
__global__ void __launch_bounds__(128)
k_cc_lift(int E, int n, const int32_t* __restrict__ cells, const uint8_t* __restrict__ cellflag,
          const double* __restrict__ dvec, const double* __restrict__ Ae, double* __restrict__ Fe) {
    constexpr int NV = D + 1, C = D + 1;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= E || !cellflag[c]) return;
    double d[NV][C];
    bool any = false;
    for (int b = 0; b < NV; ++b) {
        const int v = cells[NV * (int64_t)c + b];
        for (int l = 0; l < D; ++l) d[b][l] = dvec[D * (int64_t)v + l];
        d[b][D] = dvec[D * (int64_t)n + v];
        for (int l = 0; l < C; ++l) any = any || d[b][l] != 0.0;
    }
    if (!any) return;
    const int64_t stride = E;
    for (int a = 0; a < NV; ++a)
        for (int ri = 0; ri < C; ++ri) {
            double acc = 0.0;
            for (int b = 0; b < NV; ++b)
                for (int ci = 0; ci < C; ++ci)
                    if (d[b][ci] != 0.0) acc += Ae[(int64_t)((a * NV + b) * C * C + ri * C + ci) * stride + c] * d[b][ci];
            Fe[(int64_t)(a * C + ri) * stride + c] += acc;
        }
}

static bool cc_facet_set_active(const HemoFacetSet& fs) {
    return fs.m > 0 && (fs.coef.pconst != 0.0 || fs.coef.a_n != 0.0);
}

template <int D>
static int cc_upload_rules(hemo_ctx* ctx, const SimplexRule<D>* host_rules, const int* alias) {
    hemo_cc_state* st = ctx->cc;
    if (st->dev && st->dim == D && st->version == ctx->rule_version) return 0;
    if (st->dev && st->dim != D) { cudaFree(st->dev); st->dev = nullptr; }
    if (!st->dev) HEMO_CHECK_CUDA(ctx, cudaMalloc(&st->dev, sizeof(CcRules<D>)));
    st->dim = D;
    CcRules<D>* d = reinterpret_cast<CcRules<D>*>(st->dev);
    for (int b = 0; b < HEMO_NRULES; ++b)
        HEMO_CHECK_CUDA(ctx, cudaMemcpyAsync(&d->r[b], &host_rules[b], sizeof(SimplexRule<D>), cudaMemcpyHostToDevice, ctx->stream));
    HEMO_CHECK_CUDA(ctx, cudaMemcpyAsync(d->alias, alias, sizeof(int) * HEMO_NRULES, cudaMemcpyHostToDevice, ctx->stream));
    HEMO_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // the host copies may be temporaries
    st->version = ctx->rule_version;
    return 0;
}

template <int D>
static int cc_launch(hemo_ctx* ctx, const double* x_dev, const double* un_dev, int jac_mode, const SimplexFacetRule<D>& fr,
                     bool have_frule) {
    constexpr int NV = D + 1;
    const CcRules<D>* rules = reinterpret_cast<const CcRules<D>*>(ctx->cc->dev);
    hemo_form_finalize(ctx->par);
    const double f2 = (D == 3) ? ctx->fz : 0.0;
    dim3 grid(hemo_grid(ctx->E, 128), NV);
    k_cc_cells<D><<<grid, 128, 0, ctx->stream>>>(ctx->E, ctx->n, ctx->cells, ctx->x, ctx->h, x_dev, un_dev, ctx->par,
                                                  ctx->par.f[0], ctx->par.f[1], f2, rules, jac_mode, ctx->cellflag, ctx->Ae,
                                                  ctx->Fe);
    HEMO_LAUNCH_CHECK(ctx);
    for (int s = 0; s < HEMO_MAX_FACET_SETS; ++s) {
        const HemoFacetSet& fs = ctx->fsets[s];
        if (!cc_facet_set_active(fs)) continue;
        if (!have_frule) HEMO_FAIL(ctx, HEMO_ESTATE, "hemo_set_facet_quadrature not called");
        k_cc_facets<D><<<hemo_grid(fs.m, 128), 128, 0, ctx->stream>>>(fs.m, ctx->E, ctx->n, fs.cells, fs.mask, fs.coef, fr,
                                                                     ctx->cells, ctx->x, ctx->h, x_dev, un_dev, ctx->par,
                                                                     jac_mode, ctx->cellflag, ctx->Ae, ctx->Fe);
        HEMO_LAUNCH_CHECK(ctx);
    }
    return 0;
}

// Element tensors of the curl-curl formulation into ctx->Ae / ctx->Fe (cells, then the facet terms of every
// active set).  The caller has sized the element buffers.
int hemo_cc_cells(hemo_ctx* ctx, const double* x_dev, const double* un_dev, int jac_mode) {
    if (!ctx->have_par) HEMO_FAIL(ctx, HEMO_ESTATE, "hemo_set_params not called");
    if (ctx->par.theta != 0.5 || ctx->par.a0 != 1.0 || ctx->uh)
        HEMO_FAIL(ctx, HEMO_ESTATE, "the curl-curl formulation is written for the default time scheme (stabilized_schur_pressurebc.py:85-92)");
    if (jac_mode == CC_JAC_FLAGGED && !ctx->cellflag) jac_mode = CC_JAC_NONE;
    if (!ctx->cc) ctx->cc = new hemo_cc_state();
    int rc;
    if (ctx->dim == 3) {
        const SimplexRule<3>* r;
        const int* alias;
        const SimplexFacetRule<3>* fr;
        if ((rc = hemo_tet_get_rules(ctx, &r, &alias, &fr))) return rc;
        if ((rc = cc_upload_rules<3>(ctx, r, alias))) return rc;
        static const SimplexFacetRule<3> none{};
        return cc_launch<3>(ctx, x_dev, un_dev, jac_mode, fr ? *fr : none, fr != nullptr);
    }
    if (ctx->nv != 3) HEMO_FAIL(ctx, HEMO_ESTATE, "the curl-curl formulation is implemented on P1 triangles and tetrahedra");
    // rules of the triangle path (hemo_set_quadrature) restated as SimplexRule<2>
    if (ctx->cc->dev && ctx->cc->dim == 2 && ctx->cc->version == ctx->rule_version) {
        SimplexFacetRule<2> fr2{};
        const bool have2 = ctx->frule.nq > 0;
        if (have2) simplex_facet_rule_set<2>(fr2, ctx->frule.s, ctx->frule.w, ctx->frule.nq);
        return cc_launch<2>(ctx, x_dev, un_dev, jac_mode, fr2, have2);
    }
    std::vector<SimplexRule<2>> r(HEMO_NRULES);
    int alias[HEMO_NRULES];
    for (int b = 0; b < HEMO_NRULES; ++b) {
        if (!ctx->have_rule[b]) HEMO_FAIL(ctx, HEMO_ESTATE, "quadrature rule missing for a block form");
        const HemoRule& hr = ctx->rules[b];
        if (hr.nq > HEMO_SIMPLEX_MAXQ) return HEMO_EINVAL;
        std::vector<double> pts(2 * (size_t)hr.nq), wts((size_t)hr.nq);
        for (int q = 0; q < hr.nq; ++q) { pts[2 * q] = hr.phi[q][1]; pts[2 * q + 1] = hr.phi[q][2]; wts[q] = hr.w[q]; }
        simplex_rule_set<2>(r[b], pts.data(), wts.data(), hr.nq);
        alias[b] = hr.alias;
    }
    if ((rc = cc_upload_rules<2>(ctx, r.data(), alias))) return rc;
    SimplexFacetRule<2> fr{};
    const bool have_fr = ctx->frule.nq > 0;
    if (have_fr) {
        if (ctx->frule.nq > HEMO_SIMPLEX_MAXFQ) return HEMO_EINVAL;
        simplex_facet_rule_set<2>(fr, ctx->frule.s, ctx->frule.w, ctx->frule.nq);
    }
    return cc_launch<2>(ctx, x_dev, un_dev, jac_mode, fr, have_fr);
}

int hemo_cc_lift2d(hemo_ctx* ctx) {
    k_cc_lift<2><<<hemo_grid(ctx->E, 128), 128, 0, ctx->stream>>>(ctx->E, ctx->n, ctx->cells, ctx->cellflag, ctx->dvec, ctx->Ae,
                                                                 ctx->Fe);
    HEMO_LAUNCH_CHECK(ctx);
    return 0;
}

void hemo_cc_free(hemo_ctx* ctx) {
    if (!ctx->cc) return;
    cudaFree(ctx->cc->dev);
    delete ctx->cc;
    ctx->cc = nullptr;
}
