// Per-thread work items of the tetrahedron assembly kernels (assembly_tet.cu), written as
// __host__ __device__ functions of the thread's index so that the same code can be compiled with
// g++ and walked over an emulated grid (tests/host_simplex, tests/test_tet_host.py: test
// infrastructure only — the library never runs these on the CPU).
//
// Layouts: global vector [u interleaved (3n) | p (n)]; element buffers SoA
// Ae[(a*4+b)*16 + ri*4+ci][E], Fe[a*4+comp][E] with ri/ci/comp in (u_x, u_y, u_z, p); CSR values of
// the reference's create_matrix_block layout (src/solvers/stabilized_schur.py:191-193): with
// deg = #neighbours of node i and r0 = nrowptr[i], row 3i+k starts at 12*r0 + 4*k*deg, row 3n+i at
// 12*nnz_node + 4*r0; inside a row the u-columns of neighbour t sit at 3t..3t+2, its p-column at
// 3*deg + t.
#pragma once
#include <stdint.h>

#include "simplex_element.cuh"

// load one tetrahedron (geometry, nodal fields) and derive the fields the facet terms need
HEMO_HD void tet_load_cell(SimplexCell<3>& cd, int c, int n, const int32_t* cells, const double* x, const double* h,
                           const double* sol, const double* un, const HemoForm& par, int v[4]) {
    double X[4][3];
    for (int a = 0; a < 4; ++a) {
        v[a] = cells[4 * (int64_t)c + a];
        for (int k = 0; k < 3; ++k) {
            X[a][k] = x[3 * (int64_t)v[a] + k];
            cd.U[a][k] = sol[3 * (int64_t)v[a] + k];
            cd.N[a][k] = un[3 * (int64_t)v[a] + k];
            cd.H[a][k] = cd.N[a][k];          // the facet terms carry no time derivative
        }
        cd.P[a] = sol[3 * (int64_t)n + v[a]];
    }
    cd.fbody[0] = cd.fbody[1] = cd.fbody[2] = 0.0;
    cd.h = h[c];
    simplex_geometry<3>(cd, X);
    simplex_derive<3>(cd, par);
}

// Exterior-facet terms of boundary cell t of one tagged set: residual into Fe, derivative into Ae
// (the Jacobian assembly gathers Ae; the residual assembly needs it for the lifting of
// Dirichlet-adjacent cells, tet_lift_item).
HEMO_HD void tet_facet_item(int t, int64_t E, int n, const int32_t* fcells, const int32_t* fmask,
                            const hemo_facet_coef& co, const SimplexFacetRule<3>& fr, const int32_t* cells,
                            const double* x, const double* h, const double* sol, const double* un,
                            const HemoForm& par, bool want_jac, double* Ae, double* Fe) {
    const int c = fcells[t];
    const int mask = fmask[t];
    SimplexCell<3> cd;
    int v[4];
    tet_load_cell(cd, c, n, cells, x, h, sol, un, par, v);
    for (int lf = 0; lf < 4; ++lf) {
        if (!(mask & (1 << lf))) continue;
        auto res = [&](int a, int k, double val) { Fe[(int64_t)(a * 4 + k) * E + c] += val; };
        auto jac = [&](int a, int b, int k, int ci, double val) {
            Ae[(int64_t)((a * 4 + b) * 16 + k * 4 + ci) * E + c] += val;
        };
        if (want_jac) simplex_facet<3, true, true>(cd, par, co, fr, lf, res, jac);
        else simplex_facet<3, true, false>(cd, par, co, fr, lf, res, jac);
    }
}

// per-facet-cell partial of Q = int u_n . n ds
HEMO_HD double tet_flux_item(int t, int n, const int32_t* fcells, const int32_t* fmask, const int32_t* cells,
                             const double* x, const double* un) {
    const int c = fcells[t];
    const int mask = fmask[t];
    SimplexCell<3> cd;
    double X[4][3];
    for (int a = 0; a < 4; ++a) {
        const int v = cells[4 * (int64_t)c + a];
        for (int k = 0; k < 3; ++k) { X[a][k] = x[3 * (int64_t)v + k]; cd.N[a][k] = un[3 * (int64_t)v + k]; }
    }
    simplex_geometry<3>(cd, X);
    double q = 0.0;
    for (int lf = 0; lf < 4; ++lf)
        if (mask & (1 << lf)) q += simplex_facet_flux<3>(cd, lf);
    return q;
}

// Lifting of assemble_vector_block(..., x0 = x, alpha = -1) (stabilized_schur.py:157-175) on one
// Dirichlet-adjacent cell: Fe += Ae d with d = g - x on constrained dofs, 0 elsewhere.
HEMO_HD void tet_lift_item(int c, int64_t E, int n, const int32_t* cells, const double* dvec, const double* Ae,
                           double* Fe) {
    double d[4][4];
    bool any = false;
    for (int b = 0; b < 4; ++b) {
        const int v = cells[4 * (int64_t)c + b];
        for (int l = 0; l < 3; ++l) d[b][l] = dvec[3 * (int64_t)v + l];
        d[b][3] = dvec[3 * (int64_t)n + v];
        any = any || d[b][0] != 0.0 || d[b][1] != 0.0 || d[b][2] != 0.0 || d[b][3] != 0.0;
    }
    if (!any) return;
    for (int a = 0; a < 4; ++a)
        for (int ri = 0; ri < 4; ++ri) {
            double acc = 0.0;
            for (int b = 0; b < 4; ++b)
                for (int ci = 0; ci < 4; ++ci)
                    if (d[b][ci] != 0.0) acc += Ae[(int64_t)((a * 4 + b) * 16 + ri * 4 + ci) * E + c] * d[b][ci];
            Fe[(int64_t)(a * 4 + ri) * E + c] += acc;
        }
}

HEMO_HD int64_t tet_dof(int n, int node, int comp) {
    return comp < 3 ? 3 * (int64_t)node + comp : 3 * (int64_t)n + node;
}

// One node pair (i, j) = slot s of the node graph: sum its 4x4 contributions in the fixed order of
// the gather list, apply the Dirichlet treatment of assemble_matrix_block (rows and columns of
// constrained dofs zeroed, diagonal = number of conditions holding the dof), store 16 scalars.
HEMO_HD void tet_gather_matrix_item(int64_t s, int n, int64_t nnz_node, int64_t E, const int32_t* nrowptr,
                                    const int32_t* ncol, const int32_t* rowof, const int32_t* seg_ptr,
                                    const int32_t* seg_src, const double* Ae, const uint8_t* dofflag,
                                    const double* dofmult, double* vals) {
    const int i = rowof[s];
    double acc[16];
    for (int k = 0; k < 16; ++k) acc[k] = 0.0;
    for (int t = seg_ptr[s]; t < seg_ptr[s + 1]; ++t) {
        const int src = seg_src[t];
        const int64_t c = src / 16;
        const int ab = src - (int)c * 16;
        const double* p = Ae + (int64_t)ab * 16 * E + c;
        for (int k = 0; k < 16; ++k) acc[k] += p[k * E];
    }
    if (dofflag != nullptr) {
        const int j = ncol[s];
        bool fr[4], fc[4];
        for (int k = 0; k < 4; ++k) { fr[k] = dofflag[tet_dof(n, i, k)] != 0; fc[k] = dofflag[tet_dof(n, j, k)] != 0; }
        for (int ri = 0; ri < 4; ++ri)
            for (int ci = 0; ci < 4; ++ci)
                if (fr[ri] || fc[ci]) acc[ri * 4 + ci] = 0.0;
        if (i == j)
            for (int k = 0; k < 4; ++k)
                if (fr[k]) acc[k * 4 + k] = dofmult[tet_dof(n, i, k)];
    }
    const int r0 = nrowptr[i];
    const int deg = nrowptr[i + 1] - r0;
    const int tpos = (int)(s - r0);
    for (int ri = 0; ri < 4; ++ri) {
        const int64_t rs = (ri < 3) ? 12 * (int64_t)r0 + 4 * (int64_t)ri * deg : 12 * nnz_node + 4 * (int64_t)r0;
        vals[rs + 3 * tpos] = acc[ri * 4 + 0];
        vals[rs + 3 * tpos + 1] = acc[ri * 4 + 1];
        vals[rs + 3 * tpos + 2] = acc[ri * 4 + 2];
        vals[rs + 3 * deg + tpos] = acc[ri * 4 + 3];
    }
}

// One node: its four residual entries, then set_bc (b = x - g on constrained dofs,
// stabilized_schur.py:172-174).
HEMO_HD void tet_gather_vector_item(int i, int n, int64_t E, const int32_t* seg_ptr, const int32_t* seg_src,
                                    const double* Fe, const uint8_t* dofflag, const double* xk, const double* g,
                                    double* b) {
    double a[4] = {0, 0, 0, 0};
    for (int t = seg_ptr[i]; t < seg_ptr[i + 1]; ++t) {
        const int src = seg_src[t];
        const int64_t c = src / 4;
        const int la = src - (int)c * 4;
        for (int k = 0; k < 4; ++k) a[k] += Fe[(int64_t)(la * 4 + k) * E + c];
    }
    for (int k = 0; k < 4; ++k) {
        const int64_t d = tet_dof(n, i, k);
        if (dofflag != nullptr && dofflag[d]) a[k] = xk[d] - g[d];
        b[d] = a[k];
    }
}

// y = J x for the four rows of node i, lanes `lane`, `lane + nl`, ... of the row's neighbours
// (the kernel reduces the partial sums over the nl lanes of a node).
HEMO_HD void tet_spmv_item(int i, int lane, int nl, int n, int64_t nnz_node, const int32_t* nrowptr,
                           const int32_t* ncol, const double* vals, const double* xv, double acc[4]) {
    const int r0 = nrowptr[i];
    const int deg = nrowptr[i + 1] - r0;
    acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
    for (int t = lane; t < deg; t += nl) {
        const int j = ncol[r0 + t];
        const double x0 = xv[3 * (int64_t)j], x1 = xv[3 * (int64_t)j + 1], x2 = xv[3 * (int64_t)j + 2];
        const double x3 = xv[3 * (int64_t)n + j];
        for (int ri = 0; ri < 4; ++ri) {
            const int64_t rs = (ri < 3) ? 12 * (int64_t)r0 + 4 * (int64_t)ri * deg : 12 * nnz_node + 4 * (int64_t)r0;
            acc[ri] += vals[rs + 3 * t] * x0 + vals[rs + 3 * t + 1] * x1 + vals[rs + 3 * t + 2] * x2 +
                       vals[rs + 3 * deg + t] * x3;
        }
    }
}

// P1 stiffness and lumped mass of one tetrahedron (operators of the Schur-complement approximation,
// DESIGN.md §5): Ke[(a*4+b)][E] = |K| grad phi_a . grad phi_b, Me[a][E] = |K| / 4.
HEMO_HD void tet_laplace_item(int c, int64_t E, const int32_t* cells, const double* x, double* Ke, double* Me) {
    SimplexCell<3> cd;
    double X[4][3];
    for (int a = 0; a < 4; ++a) {
        const int v = cells[4 * (int64_t)c + a];
        for (int k = 0; k < 3; ++k) X[a][k] = x[3 * (int64_t)v + k];
    }
    simplex_geometry<3>(cd, X);
    const double vol = cd.detJ / 6.0;
    for (int a = 0; a < 4; ++a) {
        for (int b = 0; b < 4; ++b)
            Ke[(int64_t)(a * 4 + b) * E + c] = vol * (cd.g[a][0] * cd.g[b][0] + cd.g[a][1] * cd.g[b][1] + cd.g[a][2] * cd.g[b][2]);
        Me[(int64_t)a * E + c] = 0.25 * vol;
    }
}

// ---------------------------------------------------------------------------
// First 3-D preconditioner: upper block-triangular Schur factorisation like the
// 2-D path, with the velocity block approximated by a fixed number of damped block-Jacobi sweeps
// on A00 (3x3 node-diagonal blocks) read straight from the monolithic CSR values.
// ---------------------------------------------------------------------------
// inverse of the 3x3 diagonal block of node i (rows 3i..3i+2, u-columns of the node itself)
HEMO_HD void tet_dinv_item(int i, int n, int64_t nnz_node, const int32_t* nrowptr, const int32_t* diagslot,
                           const double* vals, double* dinv) {
    const int r0 = nrowptr[i];
    const int deg = nrowptr[i + 1] - r0;
    const int t = diagslot[i] - r0;
    double a[3][3];
    for (int k = 0; k < 3; ++k) {
        const int64_t rs = 12 * (int64_t)r0 + 4 * (int64_t)k * deg;
        for (int l = 0; l < 3; ++l) a[k][l] = vals[rs + 3 * t + l];
    }
    double c[3][3];
    for (int k = 0; k < 3; ++k)
        for (int l = 0; l < 3; ++l) {
            const int k1 = (k + 1) % 3, k2 = (k + 2) % 3, l1 = (l + 1) % 3, l2 = (l + 2) % 3;
            c[k][l] = a[k1][l1] * a[k2][l2] - a[k1][l2] * a[k2][l1];
        }
    const double det = a[0][0] * c[0][0] + a[0][1] * c[0][1] + a[0][2] * c[0][2];
    for (int k = 0; k < 3; ++k)
        for (int l = 0; l < 3; ++l) dinv[9 * (int64_t)i + 3 * k + l] = c[l][k] / det;
}

// t_u = r_u - A01 z_p for the three velocity rows of node i
HEMO_HD void tet_a01_residual_item(int i, int n, int64_t nnz_node, const int32_t* nrowptr, const int32_t* ncol,
                                   const double* vals, const double* zp, const double* ru, double* tu) {
    const int r0 = nrowptr[i];
    const int deg = nrowptr[i + 1] - r0;
    double acc[3] = {0.0, 0.0, 0.0};
    for (int t = 0; t < deg; ++t) {
        const double z = zp[ncol[r0 + t]];
        for (int k = 0; k < 3; ++k) acc[k] += vals[12 * (int64_t)r0 + 4 * (int64_t)k * deg + 3 * deg + t] * z;
    }
    for (int k = 0; k < 3; ++k) tu[3 * (int64_t)i + k] = ru[3 * (int64_t)i + k] - acc[k];
}

// one damped block-Jacobi sweep on A00: z_out = z_in + omega D^-1 (t - A00 z_in); z_in == nullptr: z_in = 0
HEMO_HD void tet_jacobi_item(int i, int n, int64_t nnz_node, const int32_t* nrowptr, const int32_t* ncol,
                             const double* vals, const double* dinv, double omega, const double* tu, const double* zin,
                             double* zout) {
    const int r0 = nrowptr[i];
    const int deg = nrowptr[i + 1] - r0;
    double r[3] = {tu[3 * (int64_t)i], tu[3 * (int64_t)i + 1], tu[3 * (int64_t)i + 2]};
    if (zin != nullptr) {
        for (int t = 0; t < deg; ++t) {
            const int j = ncol[r0 + t];
            const double x0 = zin[3 * (int64_t)j], x1 = zin[3 * (int64_t)j + 1], x2 = zin[3 * (int64_t)j + 2];
            for (int k = 0; k < 3; ++k) {
                const int64_t rs = 12 * (int64_t)r0 + 4 * (int64_t)k * deg + 3 * t;
                r[k] -= vals[rs] * x0 + vals[rs + 1] * x1 + vals[rs + 2] * x2;
            }
        }
    }
    for (int k = 0; k < 3; ++k) {
        const double* d = dinv + 9 * (int64_t)i + 3 * k;
        const double corr = omega * (d[0] * r[0] + d[1] * r[1] + d[2] * r[2]);
        zout[3 * (int64_t)i + k] = (zin != nullptr ? zin[3 * (int64_t)i + k] : 0.0) + corr;
    }
}

// position of `key` in the ascending range col[lo, hi), or -1
HEMO_HD int tet_find_sorted(const int32_t* col, int lo, int hi, int key) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int v = col[mid];
        if (v == key) return mid;
        if (v < key) lo = mid + 1; else hi = mid;
    }
    return -1;
}

// SELFP on tetrahedra (the reference's Schur approximation, PETSc `selfp`: stabilized_schur.py:231-235):
// entry s = (i, j) of Sp = A11 - A10 diag(A00)^-1 A01 on the distance-2 node graph (rowof2 / col2), straight
// from the monolithic CSR values in the 3-D layout.
template <typename OutT>
HEMO_HD void tet_selfp_item(int64_t s, int64_t nnz_node, const int32_t* rowof2, const int32_t* col2,
                            const int32_t* nrowptr, const int32_t* ncol, const int32_t* diagslot, const double* vals,
                            OutT* out) {
    const int i = rowof2[s], j = col2[s];
    const int r0i = nrowptr[i];
    const int degi = nrowptr[i + 1] - r0i;
    const int64_t rpi = 12 * nnz_node + 4 * (int64_t)r0i;              // pressure row of node i
    double v = 0.0;
    const int tij = tet_find_sorted(ncol, r0i, r0i + degi, j);
    if (tij >= 0) v = vals[rpi + 3 * degi + (tij - r0i)];             // A11[i, j]
    for (int t = 0; t < degi; ++t) {
        const int k = ncol[r0i + t];
        const int r0k = nrowptr[k];
        const int degk = nrowptr[k + 1] - r0k;
        const int tkj = tet_find_sorted(ncol, r0k, r0k + degk, j);
        if (tkj < 0) continue;
        const int tkk = diagslot[k] - r0k;
        for (int c = 0; c < 3; ++c) {
            const int64_t ru = 12 * (int64_t)r0k + 4 * (int64_t)c * degk;  // row 3k + c
            const double a10 = vals[rpi + 3 * t + c];                  // A10[i, (k, c)]
            const double a01 = vals[ru + 3 * degk + (tkj - r0k)];      // A01[(k, c), j]
            const double d = vals[ru + 3 * tkk + c];                   // diag(A00)
            v -= a10 * a01 / d;
        }
    }
    out[s] = (OutT)v;
}

// ---------------------------------------------------------------------------
// Per-step post-processing on tetrahedra (postproc.cu)
// ---------------------------------------------------------------------------
// int_K |f|^2 for a P1 field with bs interleaved components: |J| / 120 * sum_k ((sum_a F_a)^2 + sum_a F_a^2)
HEMO_HD double tet_l2_item(int c, int bs, const int32_t* cells, const double* x, const double* f) {
    SimplexCell<3> cd;
    double X[4][3];
    int v[4];
    for (int a = 0; a < 4; ++a) {
        v[a] = cells[4 * (int64_t)c + a];
        for (int k = 0; k < 3; ++k) X[a][k] = x[3 * (int64_t)v[a] + k];
    }
    simplex_geometry<3>(cd, X);
    double acc = 0.0;
    for (int k = 0; k < bs; ++k) {
        double s1 = 0.0, s2 = 0.0;
        for (int a = 0; a < 4; ++a) {
            const double F = f[(int64_t)v[a] * bs + k];
            s1 += F; s2 += F * F;
        }
        acc += s1 * s1 + s2;
    }
    return cd.detJ / 120.0 * acc;
}

// solver.assemble_wss() (src/solverBase.py:144-195) on the tagged facets of boundary cell t:
// add(node, k, value) receives (1/|F|) int_F phi_i ds * Tt_k = Tt_k / 3 for the three facet vertices,
// Tt = T - (T.n) n, T = -2 mu eps(u) n (eps constant on a P1 cell).
template <typename Add>
HEMO_HD void tet_wss_item(int t, const int32_t* fcells, const int32_t* fmask, const int32_t* cells, const double* x,
                          const double* sol, double mu, Add add) {
    const int c = fcells[t];
    const int mask = fmask[t];
    SimplexCell<3> cd;
    double X[4][3], U[4][3];
    int v[4];
    for (int a = 0; a < 4; ++a) {
        v[a] = cells[4 * (int64_t)c + a];
        for (int k = 0; k < 3; ++k) { X[a][k] = x[3 * (int64_t)v[a] + k]; U[a][k] = sol[3 * (int64_t)v[a] + k]; }
    }
    simplex_geometry<3>(cd, X);
    double G[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double g = 0.0;
            for (int a = 0; a < 4; ++a) g += cd.g[a][i] * U[a][j];
            G[i][j] = g;
        }
    for (int lf = 0; lf < 4; ++lf) {
        if (!(mask & (1 << lf))) continue;
        double nr[3], scale;
        simplex_facet_normal<3>(cd, lf, nr, scale);
        double T[3], tn = 0.0;
        for (int i = 0; i < 3; ++i) {
            double e = 0.0;
            for (int j = 0; j < 3; ++j) e += 0.5 * (G[i][j] + G[j][i]) * nr[j];
            T[i] = -2.0 * mu * e;
            tn += T[i] * nr[i];
        }
        for (int a = 0; a < 4; ++a) {
            if (a == lf) continue;
            for (int k = 0; k < 3; ++k) add(v[a], k, (T[k] - tn * nr[k]) / 3.0);
        }
    }
}
