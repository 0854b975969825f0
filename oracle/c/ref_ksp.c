/* CPU ORACLE / BASELINE (C + OpenMP) — TEST AND BENCHMARK INFRASTRUCTURE ONLY, never linked into the product.
 *
 * Restates, for the host cores, the PETSc solver configuration the reference sets up in
 *   /root/reference/src/solvers/stabilized_schur.py:226-275
 *     KSP fgmres (right PC, restart 200, rtol 1e-5, max_it 1000, classical Gram-Schmidt)
 *     PC  fieldsplit SCHUR, factorisation FULL, Schur preconditioner SELFP
 *         (Sp = A11 - A10 diag(A00)^-1 A01, re-formed for every Jacobian, :235,253)
 *     sub-KSP u: gmres(30) (left PC, rtol 1e-5, max_it 10000) + asm (overlap 1, restricted, ILU(0) per block)
 *     sub-KSP p: preonly + asm (ILU(0) per block) on Sp                                   (:256-267)
 *   and the constant-pressure null-space projection of :283-293,314-319.
 * PETSc itself is a third-party dependency that is not installable here: this is a restatement
 * of its published algorithms (Saad's FGMRES / GMRES, ILU(0) in natural ordering, restricted additive Schwarz with
 * one block per MPI rank), with one OpenMP thread standing for one MPI rank: the ASM blocks are contiguous
 * row ranges (what DOLFINx's ownership ranges are), vector operations and SpMV are split the same way.
 * PARITY UNPINNED against PETSc; the converged solution is compared with the sparse-LU oracle in
 * tests/test_cpu_reference.py.
 *
 * Also here: the threaded insertion of element tensors into the fixed CSR pattern (MatSetValuesLocal(ADD),
 * stabilized_schur.py:154) used by the CPU arm of bench.py.
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n;
    int64_t* rp;
    int* ci;
    double* v;
} csr_t;

static void csr_free(csr_t* A) {
    free(A->rp); free(A->ci); free(A->v);
    A->rp = NULL; A->ci = NULL; A->v = NULL; A->n = 0;
}

/* ---------------------------------------------------------------------------------------------
 * vector kernels (static schedule = the same contiguous ownership ranges everywhere)
 * ------------------------------------------------------------------------------------------- */
static double vdot(int64_t n, const double* x, const double* y) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int64_t i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}
static double vnorm(int64_t n, const double* x) { return sqrt(vdot(n, x, x)); }
static void vaxpy(int64_t n, double a, const double* x, double* y) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) y[i] += a * x[i];
}
static void vscale_copy(int64_t n, double a, const double* x, double* y) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) y[i] = a * x[i];
}
static void vzero(int64_t n, double* x) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) x[i] = 0.0;
}
static void spmv(const csr_t* A, const double* x, double* y) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < A->n; ++i) {
        double s = 0.0;
        for (int64_t t = A->rp[i]; t < A->rp[i + 1]; ++t) s += A->v[t] * x[A->ci[t]];
        y[i] = s;
    }
}
/* h[i] = V_i . w, i < k (VecMDot): one pass over w per thread chunk */
static void vmdot(int64_t n, int k, double* const* V, const double* w, double* h) {
    for (int i = 0; i < k; ++i) h[i] = 0.0;
#pragma omp parallel
    {
        double* loc = (double*)calloc((size_t)k, sizeof(double));
#pragma omp for schedule(static) nowait
        for (int64_t b = 0; b < (n + 1023) / 1024; ++b) {
            const int64_t i0 = b * 1024, i1 = i0 + 1024 < n ? i0 + 1024 : n;
            for (int i = 0; i < k; ++i) {
                const double* v = V[i];
                double s = 0.0;
                for (int64_t q = i0; q < i1; ++q) s += v[q] * w[q];
                loc[i] += s;
            }
        }
#pragma omp critical
        for (int i = 0; i < k; ++i) h[i] += loc[i];
        free(loc);
    }
}
/* w += sign * sum_i c[i] V_i (VecMAXPY) */
static void vmaxpy(int64_t n, int k, double* const* V, const double* c, double sign, double* w) {
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < (n + 1023) / 1024; ++b) {
        const int64_t i0 = b * 1024, i1 = i0 + 1024 < n ? i0 + 1024 : n;
        for (int i = 0; i < k; ++i) {
            const double a = sign * c[i];
            const double* v = V[i];
            for (int64_t q = i0; q < i1; ++q) w[q] += a * v[q];
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * PCASM (restricted, overlap 1) with ILU(0) blocks: one block per "rank"
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int lo, hi;    /* owned rows */
    int m;         /* rows of the overlapping block */
    int o0;        /* local index of row lo */
    int* ext;      /* m sorted global rows */
    int64_t* rp;   /* local CSR pattern (columns = local indices, sorted) */
    int* ci;
    int64_t* src;  /* position of each local entry in the global value array */
    int64_t* diag; /* position of the diagonal in each local row */
    double* lu;    /* ILU(0) factors, same pattern */
    double* wrk;   /* m */
} asm_block_t;

typedef struct {
    int nb;
    asm_block_t* b;
} asm_t;

static int cmp_int(const void* a, const void* b) { return (*(const int*)a > *(const int*)b) - (*(const int*)a < *(const int*)b); }

static int find_sorted(const int* a, int n, int key) {
    int lo = 0, hi = n - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] == key) return mid;
        if (a[mid] < key) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

static void asm_free(asm_t* P) {
    for (int t = 0; t < P->nb; ++t) {
        asm_block_t* B = &P->b[t];
        free(B->ext); free(B->rp); free(B->ci); free(B->src); free(B->diag); free(B->lu); free(B->wrk);
    }
    free(P->b);
    P->b = NULL; P->nb = 0;
}

static void asm_symbolic(asm_t* P, const csr_t* A, int nb) {
    P->nb = nb;
    P->b = (asm_block_t*)calloc((size_t)nb, sizeof(asm_block_t));
    const int n = A->n;
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nb; ++t) {
        asm_block_t* B = &P->b[t];
        B->lo = (int)((int64_t)n * t / nb);
        B->hi = (int)((int64_t)n * (t + 1) / nb);
        /* overlap 1: owned rows plus every column they touch */
        const int64_t cnt = A->rp[B->hi] - A->rp[B->lo];
        int* tmp = (int*)malloc(sizeof(int) * (size_t)(cnt + (B->hi - B->lo) + 1));
        int64_t k = 0;
        for (int i = B->lo; i < B->hi; ++i) tmp[k++] = i;
        for (int64_t q = A->rp[B->lo]; q < A->rp[B->hi]; ++q) {
            const int c = A->ci[q];
            if (c < B->lo || c >= B->hi) tmp[k++] = c;
        }
        qsort(tmp, (size_t)k, sizeof(int), cmp_int);
        int m = 0;
        for (int64_t q = 0; q < k; ++q)
            if (m == 0 || tmp[q] != tmp[m - 1]) tmp[m++] = tmp[q];
        B->m = m;
        B->ext = (int*)malloc(sizeof(int) * (size_t)m);
        memcpy(B->ext, tmp, sizeof(int) * (size_t)m);
        free(tmp);
        B->o0 = find_sorted(B->ext, m, B->lo);
        B->rp = (int64_t*)malloc(sizeof(int64_t) * (size_t)(m + 1));
        int64_t nnz = 0;
        for (int li = 0; li < m; ++li) {
            const int gi = B->ext[li];
            B->rp[li] = nnz;
            for (int64_t q = A->rp[gi]; q < A->rp[gi + 1]; ++q)
                if (find_sorted(B->ext, m, A->ci[q]) >= 0) ++nnz;
        }
        B->rp[m] = nnz;
        B->ci = (int*)malloc(sizeof(int) * (size_t)nnz);
        B->src = (int64_t*)malloc(sizeof(int64_t) * (size_t)nnz);
        B->lu = (double*)malloc(sizeof(double) * (size_t)nnz);
        B->diag = (int64_t*)malloc(sizeof(int64_t) * (size_t)m);
        B->wrk = (double*)malloc(sizeof(double) * (size_t)m);
        for (int li = 0; li < m; ++li) {
            const int gi = B->ext[li];
            int64_t p = B->rp[li];
            B->diag[li] = -1;
            for (int64_t q = A->rp[gi]; q < A->rp[gi + 1]; ++q) {
                const int lc = find_sorted(B->ext, m, A->ci[q]);
                if (lc < 0) continue;
                B->ci[p] = lc;
                B->src[p] = q;
                if (lc == li) B->diag[li] = p;
                ++p;
            }
        }
    }
}

/* ILU(0), natural ordering, IKJ variant (PETSc MatILUFactor levels = 0) */
static void asm_numeric(asm_t* P, const csr_t* A) {
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < P->nb; ++t) {
        asm_block_t* B = &P->b[t];
        const int m = B->m;
        for (int64_t p = 0; p < B->rp[m]; ++p) B->lu[p] = A->v[B->src[p]];
        int64_t* pos = (int64_t*)malloc(sizeof(int64_t) * (size_t)m);
        for (int i = 0; i < m; ++i) pos[i] = -1;
        for (int i = 0; i < m; ++i) {
            for (int64_t p = B->rp[i]; p < B->rp[i + 1]; ++p) pos[B->ci[p]] = p;
            for (int64_t p = B->rp[i]; p < B->rp[i + 1]; ++p) {
                const int k = B->ci[p];
                if (k >= i) break;
                double dk = B->lu[B->diag[k]];
                if (dk == 0.0) dk = 1e-300;
                const double lik = B->lu[p] / dk;
                B->lu[p] = lik;
                for (int64_t q = B->diag[k] + 1; q < B->rp[k + 1]; ++q) {
                    const int64_t w = pos[B->ci[q]];
                    if (w >= 0) B->lu[w] -= lik * B->lu[q];
                }
            }
            for (int64_t p = B->rp[i]; p < B->rp[i + 1]; ++p) pos[B->ci[p]] = -1;
        }
        free(pos);
    }
}

/* z = sum_blocks R0_t^T (L_t U_t)^-1 R_t r   (restricted ASM: only the owned rows of each block are kept) */
static void asm_apply(const asm_t* P, const double* r, double* z) {
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < P->nb; ++t) {
        const asm_block_t* B = &P->b[t];
        const int m = B->m;
        double* y = B->wrk;
        for (int i = 0; i < m; ++i) {
            double s = r[B->ext[i]];
            for (int64_t p = B->rp[i]; p < B->diag[i]; ++p) s -= B->lu[p] * y[B->ci[p]];
            y[i] = s;
        }
        for (int i = m - 1; i >= 0; --i) {
            double s = y[i];
            for (int64_t p = B->diag[i] + 1; p < B->rp[i + 1]; ++p) s -= B->lu[p] * y[B->ci[p]];
            double d = B->lu[B->diag[i]];
            if (d == 0.0) d = 1e-300;
            y[i] = s / d;
        }
        for (int i = B->lo; i < B->hi; ++i) z[i] = y[B->o0 + (i - B->lo)];
    }
}

/* ---------------------------------------------------------------------------------------------
 * the solver object
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int nu, np, nranks;
    int nullspace;
    csr_t A;                       /* borrowed arrays (rp, ci, v) of the monolithic Jacobian */
    csr_t A00, A01, A10, A11, Sp;
    int64_t *s00, *s01, *s10, *s11; /* source positions of the sub-block entries */
    asm_t pcu, pcp;
    int have_symbolic;
    /* inner GMRES(30) workspace */
    double* gv[32];
    double *gw, *gt;
    /* FGMRES workspace (allocated on demand, PETSc allocates its basis in chunks too) */
    int restart;
    double **V, **Z;
    double *w, *ru, *rp_, *zu, *zp, *tu, *tp;
    /* statistics */
    int64_t inner_its, inner_solves, outer_its;
    int inner_restart, inner_maxit;
    double inner_rtol;
} refksp_t;

refksp_t* refksp_create(int nu, int np, int nranks, int restart) {
    refksp_t* S = (refksp_t*)calloc(1, sizeof(refksp_t));
    S->nu = nu; S->np = np; S->nranks = nranks > 0 ? nranks : omp_get_max_threads();
    S->restart = restart > 0 ? restart : 200;
    S->inner_restart = 30; S->inner_maxit = 10000; S->inner_rtol = 1e-5;
    const int64_t N = (int64_t)nu + np;
    for (int i = 0; i < 32; ++i) S->gv[i] = (double*)malloc(sizeof(double) * (size_t)nu);
    S->gw = (double*)malloc(sizeof(double) * (size_t)nu);
    S->gt = (double*)malloc(sizeof(double) * (size_t)nu);
    S->V = (double**)calloc((size_t)S->restart + 1, sizeof(double*));
    S->Z = (double**)calloc((size_t)S->restart, sizeof(double*));
    S->w = (double*)malloc(sizeof(double) * (size_t)N);
    S->ru = (double*)malloc(sizeof(double) * (size_t)nu);
    S->zu = (double*)malloc(sizeof(double) * (size_t)nu);
    S->tu = (double*)malloc(sizeof(double) * (size_t)nu);
    S->rp_ = (double*)malloc(sizeof(double) * (size_t)np);
    S->zp = (double*)malloc(sizeof(double) * (size_t)np);
    S->tp = (double*)malloc(sizeof(double) * (size_t)np);
    return S;
}

void refksp_destroy(refksp_t* S) {
    if (!S) return;
    csr_free(&S->A00); csr_free(&S->A01); csr_free(&S->A10); csr_free(&S->A11); csr_free(&S->Sp);
    free(S->s00); free(S->s01); free(S->s10); free(S->s11);
    if (S->have_symbolic) { asm_free(&S->pcu); asm_free(&S->pcp); }
    for (int i = 0; i < 32; ++i) free(S->gv[i]);
    free(S->gw); free(S->gt);
    for (int i = 0; i <= S->restart; ++i) free(S->V[i]);
    for (int i = 0; i < S->restart; ++i) free(S->Z[i]);
    free(S->V); free(S->Z); free(S->w); free(S->ru); free(S->zu); free(S->tu); free(S->rp_); free(S->zp); free(S->tp);
    free(S);
}

void refksp_set_inner(refksp_t* S, int restart, int maxit, double rtol) {
    S->inner_restart = restart < 31 ? restart : 30; S->inner_maxit = maxit; S->inner_rtol = rtol;
}
void refksp_set_nullspace(refksp_t* S, int on) { S->nullspace = on; }
void refksp_stats(refksp_t* S, int64_t* out) { out[0] = S->outer_its; out[1] = S->inner_its; out[2] = S->inner_solves; }

/* MatCreateSubMatrix for the four blocks (PCSetUp_FieldSplit): patterns once, values every time */
static void split_symbolic(refksp_t* S) {
    const csr_t* A = &S->A;
    const int nu = S->nu, np = S->np;
    csr_t* blk[4] = {&S->A00, &S->A01, &S->A10, &S->A11};
    int64_t** src[4] = {&S->s00, &S->s01, &S->s10, &S->s11};
    for (int b = 0; b < 4; ++b) {
        const int r0 = (b < 2) ? 0 : nu, nr = (b < 2) ? nu : np;
        const int c0 = (b & 1) ? nu : 0, c1 = (b & 1) ? nu + np : nu;
        csr_t* B = blk[b];
        B->n = nr;
        B->rp = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nr + 1));
        B->rp[0] = 0;
        for (int i = 0; i < nr; ++i) {
            int64_t c = 0;
            for (int64_t q = A->rp[r0 + i]; q < A->rp[r0 + i + 1]; ++q) c += (A->ci[q] >= c0 && A->ci[q] < c1);
            B->rp[i + 1] = B->rp[i] + c;
        }
        const int64_t nnz = B->rp[nr];
        B->ci = (int*)malloc(sizeof(int) * (size_t)nnz);
        B->v = (double*)malloc(sizeof(double) * (size_t)nnz);
        *src[b] = (int64_t*)malloc(sizeof(int64_t) * (size_t)nnz);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < nr; ++i) {
            int64_t p = B->rp[i];
            for (int64_t q = A->rp[r0 + i]; q < A->rp[r0 + i + 1]; ++q)
                if (A->ci[q] >= c0 && A->ci[q] < c1) { B->ci[p] = A->ci[q] - c0; (*src[b])[p] = q; ++p; }
        }
    }
}

static void split_numeric(refksp_t* S) {
    csr_t* blk[4] = {&S->A00, &S->A01, &S->A10, &S->A11};
    int64_t* src[4] = {S->s00, S->s01, S->s10, S->s11};
    for (int b = 0; b < 4; ++b) {
        csr_t* B = blk[b];
        const int64_t nnz = B->rp[B->n];
#pragma omp parallel for schedule(static)
        for (int64_t p = 0; p < nnz; ++p) B->v[p] = S->A.v[src[b][p]];
    }
}

/* Sp = A11 - A10 diag(A00)^-1 A01 (MatSchurComplementGetPmat, SELFP): row-wise Gustavson product */
static void selfp(refksp_t* S, int symbolic) {
    const int np = S->np, nu = S->nu;
    double* dinv = S->gt;     /* nu scratch */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nu; ++i) {
        double d = 0.0;
        for (int64_t q = S->A00.rp[i]; q < S->A00.rp[i + 1]; ++q)
            if (S->A00.ci[q] == i) d = S->A00.v[q];
        dinv[i] = d != 0.0 ? 1.0 / d : 0.0;
    }
    csr_t* Sp = &S->Sp;
    if (symbolic) {
        Sp->n = np;
        Sp->rp = (int64_t*)calloc((size_t)np + 1, sizeof(int64_t));
        int** rows = (int**)calloc((size_t)np, sizeof(int*));
#pragma omp parallel
        {
            int* mark = (int*)malloc(sizeof(int) * (size_t)np);
            int* list = (int*)malloc(sizeof(int) * (size_t)np);
            for (int i = 0; i < np; ++i) mark[i] = -1;
#pragma omp for schedule(static)
            for (int i = 0; i < np; ++i) {
                int cnt = 0;
                for (int64_t q = S->A11.rp[i]; q < S->A11.rp[i + 1]; ++q) {
                    const int c = S->A11.ci[q];
                    if (mark[c] != i) { mark[c] = i; list[cnt++] = c; }
                }
                for (int64_t q = S->A10.rp[i]; q < S->A10.rp[i + 1]; ++q) {
                    const int k = S->A10.ci[q];
                    for (int64_t r = S->A01.rp[k]; r < S->A01.rp[k + 1]; ++r) {
                        const int c = S->A01.ci[r];
                        if (mark[c] != i) { mark[c] = i; list[cnt++] = c; }
                    }
                }
                qsort(list, (size_t)cnt, sizeof(int), cmp_int);
                rows[i] = (int*)malloc(sizeof(int) * (size_t)(cnt > 0 ? cnt : 1));
                memcpy(rows[i], list, sizeof(int) * (size_t)cnt);
                Sp->rp[i + 1] = cnt;
            }
            free(mark); free(list);
        }
        for (int i = 0; i < np; ++i) Sp->rp[i + 1] += Sp->rp[i];
        Sp->ci = (int*)malloc(sizeof(int) * (size_t)Sp->rp[np]);
        Sp->v = (double*)malloc(sizeof(double) * (size_t)Sp->rp[np]);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < np; ++i) {
            memcpy(Sp->ci + Sp->rp[i], rows[i], sizeof(int) * (size_t)(Sp->rp[i + 1] - Sp->rp[i]));
            free(rows[i]);
        }
        free(rows);
    }
#pragma omp parallel
    {
        double* acc = (double*)calloc((size_t)np, sizeof(double));
#pragma omp for schedule(static)
        for (int i = 0; i < np; ++i) {
            for (int64_t q = S->A11.rp[i]; q < S->A11.rp[i + 1]; ++q) acc[S->A11.ci[q]] += S->A11.v[q];
            for (int64_t q = S->A10.rp[i]; q < S->A10.rp[i + 1]; ++q) {
                const int k = S->A10.ci[q];
                const double a = S->A10.v[q] * dinv[k];
                if (a == 0.0) continue;
                for (int64_t r = S->A01.rp[k]; r < S->A01.rp[k + 1]; ++r) acc[S->A01.ci[r]] -= a * S->A01.v[r];
            }
            for (int64_t p = Sp->rp[i]; p < Sp->rp[i + 1]; ++p) { Sp->v[p] = acc[Sp->ci[p]]; acc[Sp->ci[p]] = 0.0; }
        }
        free(acc);
    }
}

/* KSPSetOperators + PCSetUp: called for every new Jacobian (rowptr/colind/vals are borrowed) */
void refksp_setup(refksp_t* S, int64_t* rowptr, int* colind, double* vals) {
    S->A.n = S->nu + S->np; S->A.rp = rowptr; S->A.ci = colind; S->A.v = vals;
    omp_set_num_threads(S->nranks);
    const int first = !S->have_symbolic;
    if (first) split_symbolic(S);
    split_numeric(S);
    selfp(S, first);
    if (first) {
        asm_symbolic(&S->pcu, &S->A00, S->nranks);
        asm_symbolic(&S->pcp, &S->Sp, S->nranks);
        S->have_symbolic = 1;
    }
    asm_numeric(&S->pcu, &S->A00);
    asm_numeric(&S->pcp, &S->Sp);
}

/* sub-KSP u: left-preconditioned GMRES(30) on A00, zero initial guess, rtol on the preconditioned residual */
static void inner_gmres(refksp_t* S, const double* b, double* x) {
    const int n = S->nu, m = S->inner_restart;
    double H[31][30], cs[30], sn[30], g[31], y[30];
    vzero(n, x);
    asm_apply(&S->pcu, b, S->gv[0]);              /* r0 = M^-1 b */
    double beta = vnorm(n, S->gv[0]);
    const double tol = S->inner_rtol * beta;
    int its = 0;
    S->inner_solves++;
    if (beta == 0.0) return;
    while (its < S->inner_maxit) {
        vscale_copy(n, 1.0 / beta, S->gv[0], S->gv[0]);
        memset(g, 0, sizeof g);
        g[0] = beta;
        int j = 0, done = 0;
        for (; j < m && its < S->inner_maxit; ++j) {
            spmv(&S->A00, S->gv[j], S->gt);
            asm_apply(&S->pcu, S->gt, S->gw);      /* w = M^-1 A v_j */
            double h[31];
            vmdot(n, j + 1, S->gv, S->gw, h);
            vmaxpy(n, j + 1, S->gv, h, -1.0, S->gw);
            const double hn = vnorm(n, S->gw);
            for (int i = 0; i <= j; ++i) H[i][j] = h[i];
            H[j + 1][j] = hn;
            if (hn > 0.0) vscale_copy(n, 1.0 / hn, S->gw, S->gv[j + 1]);
            for (int i = 0; i < j; ++i) {
                const double a = H[i][j], c = H[i + 1][j];
                H[i][j] = cs[i] * a + sn[i] * c;
                H[i + 1][j] = -sn[i] * a + cs[i] * c;
            }
            const double d = hypot(H[j][j], H[j + 1][j]);
            cs[j] = d > 0 ? H[j][j] / d : 1.0;
            sn[j] = d > 0 ? H[j + 1][j] / d : 0.0;
            H[j][j] = d; H[j + 1][j] = 0.0;
            g[j + 1] = -sn[j] * g[j];
            g[j] = cs[j] * g[j];
            ++its;
            if (fabs(g[j + 1]) <= tol || hn == 0.0) { done = 1; ++j; break; }
        }
        const int k = j;
        for (int i = k - 1; i >= 0; --i) {
            double s = g[i];
            for (int l = i + 1; l < k; ++l) s -= H[i][l] * y[l];
            y[i] = s / H[i][i];
        }
        vmaxpy(n, k, S->gv, y, 1.0, x);
        if (done) break;
        /* restart: r = M^-1 (b - A x) */
        spmv(&S->A00, x, S->gt);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; ++i) S->gt[i] = b[i] - S->gt[i];
        asm_apply(&S->pcu, S->gt, S->gv[0]);
        beta = vnorm(n, S->gv[0]);
        if (beta <= tol) break;
    }
    S->inner_its += its;
}

static void remove_mean(int n, double* x) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int i = 0; i < n; ++i) s += x[i];
    s /= (double)n;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) x[i] -= s;
}

/* PCApply_FieldSplit_Schur, FULL factorisation:
 *   z_u = A00^-1 r_u;  z_p = Sp^-1 (r_p - A10 z_u);  z_u = A00^-1 (r_u - A01 z_p)            */
static void pc_apply(refksp_t* S, const double* r, double* z) {
    const int nu = S->nu, np = S->np;
    inner_gmres(S, r, S->zu);
    spmv(&S->A10, S->zu, S->tp);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < np; ++i) S->tp[i] = r[nu + i] - S->tp[i];
    if (S->nullspace) remove_mean(np, S->tp);
    asm_apply(&S->pcp, S->tp, z + nu);
    if (S->nullspace) remove_mean(np, z + nu);
    spmv(&S->A01, z + nu, S->tu);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nu; ++i) S->tu[i] = r[i] - S->tu[i];
    inner_gmres(S, S->tu, z);
}

/* KSPSolve(fgmres): right-preconditioned, zero initial guess.  Returns 0 converged, 1 max_it. */
int refksp_solve(refksp_t* S, const double* b, double* x, double rtol, int max_it, int* its_out, double* rel_out) {
    omp_set_num_threads(S->nranks);
    const int64_t N = (int64_t)S->nu + S->np;
    const int m = S->restart;
    double* H = (double*)calloc((size_t)(m + 1) * m, sizeof(double));
    double *cs = (double*)calloc((size_t)m, sizeof(double)), *sn = (double*)calloc((size_t)m, sizeof(double));
    double *g = (double*)calloc((size_t)m + 1, sizeof(double)), *y = (double*)calloc((size_t)m, sizeof(double));
    double* h = (double*)calloc((size_t)m + 1, sizeof(double));
    vzero(N, x);
    const double bnorm = vnorm(N, b);
    int its = 0, converged = 0;
    double res = bnorm;
    if (bnorm == 0.0) { converged = 1; goto out; }
    const double tol = rtol * bnorm;
    double beta = bnorm;
    if (!S->V[0]) S->V[0] = (double*)malloc(sizeof(double) * (size_t)N);
    vscale_copy(N, 1.0 / beta, b, S->V[0]);
    while (!converged && its < max_it) {
        for (int i = 0; i <= m; ++i) g[i] = 0.0;
        g[0] = beta;
        int j = 0;
        for (; j < m && its < max_it; ++j) {
            if (!S->Z[j]) S->Z[j] = (double*)malloc(sizeof(double) * (size_t)N);
            if (!S->V[j + 1]) S->V[j + 1] = (double*)malloc(sizeof(double) * (size_t)N);
            pc_apply(S, S->V[j], S->Z[j]);
            spmv(&S->A, S->Z[j], S->w);
            vmdot(N, j + 1, S->V, S->w, h);
            vmaxpy(N, j + 1, S->V, h, -1.0, S->w);
            const double hn = vnorm(N, S->w);
            for (int i = 0; i <= j; ++i) H[(size_t)i * m + j] = h[i];
            H[(size_t)(j + 1) * m + j] = hn;
            if (hn > 0.0) vscale_copy(N, 1.0 / hn, S->w, S->V[j + 1]);
            for (int i = 0; i < j; ++i) {
                const double a = H[(size_t)i * m + j], c = H[(size_t)(i + 1) * m + j];
                H[(size_t)i * m + j] = cs[i] * a + sn[i] * c;
                H[(size_t)(i + 1) * m + j] = -sn[i] * a + cs[i] * c;
            }
            const double d = hypot(H[(size_t)j * m + j], H[(size_t)(j + 1) * m + j]);
            cs[j] = d > 0 ? H[(size_t)j * m + j] / d : 1.0;
            sn[j] = d > 0 ? H[(size_t)(j + 1) * m + j] / d : 0.0;
            H[(size_t)j * m + j] = d; H[(size_t)(j + 1) * m + j] = 0.0;
            g[j + 1] = -sn[j] * g[j];
            g[j] = cs[j] * g[j];
            ++its;
            res = fabs(g[j + 1]);
            if (!isfinite(res)) { j++; goto out; }
            if (res <= tol || hn == 0.0) { converged = 1; ++j; break; }
        }
        const int k = j;
        for (int i = k - 1; i >= 0; --i) {
            double s = g[i];
            for (int l = i + 1; l < k; ++l) s -= H[(size_t)i * m + l] * y[l];
            y[i] = s / H[(size_t)i * m + i];
        }
        vmaxpy(N, k, S->Z, y, 1.0, x);
        if (converged) break;
        spmv(&S->A, x, S->w);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i) S->w[i] = b[i] - S->w[i];
        beta = vnorm(N, S->w);
        res = beta;
        if (beta <= tol) { converged = 1; break; }
        vscale_copy(N, 1.0 / beta, S->w, S->V[0]);
    }
out:
    S->outer_its += its;
    if (its_out) *its_out = its;
    if (rel_out) *rel_out = bnorm > 0 ? res / bnorm : 0.0;
    free(H); free(cs); free(sn); free(g); free(y); free(h);
    return converged ? 0 : 1;
}

/* ---------------------------------------------------------------------------------------------
 * MatSetValuesLocal(ADD_VALUES) of the element tensors into the fixed CSR pattern, threaded over
 * cells.  pos[e*nl*nl + r*nl + c] = CSR position of local entry (r, c) of cell e.
 * ------------------------------------------------------------------------------------------- */
void ref_insert_matrix(int64_t E, int nl, const double* Ae, const int32_t* pos, int64_t nnz, double* vals, int nthreads) {
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < nnz; ++q) vals[q] = 0.0;
    const int64_t per = (int64_t)nl * nl;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < E; ++e) {
        const double* a = Ae + e * per;
        const int32_t* p = pos + e * per;
        for (int64_t k = 0; k < per; ++k) {
#pragma omp atomic update
            vals[p[k]] += a[k];
        }
    }
}

void ref_insert_vector(int64_t E, int nl, const double* Fe, const int64_t* l2g, int64_t N, double* b, int nthreads) {
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < N; ++q) b[q] = 0.0;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < E; ++e) {
        for (int k = 0; k < nl; ++k) {
#pragma omp atomic update
            b[l2g[e * nl + k]] += Fe[e * nl + k];
        }
    }
}

/* positions of the element entries in the CSR pattern (binary search per entry; one-time) */
void ref_cell_positions(int64_t E, int nl, const int64_t* l2g, const int64_t* rowptr, const int32_t* colind, int32_t* pos,
                        int nthreads) {
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < E; ++e) {
        for (int r = 0; r < nl; ++r) {
            const int64_t row = l2g[e * nl + r];
            const int64_t r0 = rowptr[row], r1 = rowptr[row + 1];
            for (int c = 0; c < nl; ++c) {
                const int key = (int)l2g[e * nl + c];
                int64_t lo = r0, hi = r1 - 1, f = -1;
                while (lo <= hi) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (colind[mid] == key) { f = mid; break; }
                    if (colind[mid] < key) lo = mid + 1; else hi = mid - 1;
                }
                pos[(e * nl + r) * nl + c] = (int32_t)f;
            }
        }
    }
}
