// Host-side (CPU) one-time setup helpers exported through the C-ABI.
// These run once per mesh, like DOLFINx's dofmap / sparsity-pattern builders
// run once inside `create_matrix_block` (reference src/solvers/stabilized_schur.py:191);
// they are not on the per-timestep path.
#include <vector>

#include "hemo_internal.cuh"

// Greedy (Vanek-style) aggregation on a strength graph given as CSR without
// diagonal.  exclude[i] != 0 removes node i (Dirichlet nodes): agg[i] = -1.
// Isolated, non-excluded nodes become singleton aggregates.
extern "C" int hemo_host_aggregate(int n, const int32_t* rowptr, const int32_t* col, const uint8_t* exclude,
                                   int32_t* agg, int32_t* n_agg_out) {
    if (n <= 0 || !rowptr || !col || !agg || !n_agg_out) return HEMO_EINVAL;
    for (int i = 0; i < n; ++i) agg[i] = -1;
    auto excluded = [&](int i) { return exclude && exclude[i]; };
    int na = 0;
    // pass 1: roots whose whole strong neighbourhood is free
    for (int i = 0; i < n; ++i) {
        if (agg[i] != -1 || excluded(i)) continue;
        bool free_nb = true;
        int cnt = 0;
        for (int t = rowptr[i]; t < rowptr[i + 1]; ++t) {
            const int j = col[t];
            if (j == i || excluded(j)) continue;
            ++cnt;
            if (agg[j] != -1) { free_nb = false; break; }
        }
        if (!free_nb || cnt == 0) continue;
        agg[i] = na;
        for (int t = rowptr[i]; t < rowptr[i + 1]; ++t) {
            const int j = col[t];
            if (j != i && !excluded(j)) agg[j] = na;
        }
        ++na;
    }
    // pass 2: attach leftovers to a neighbouring pass-1 aggregate
    std::vector<int32_t> tent(n, -1);
    for (int i = 0; i < n; ++i) {
        if (agg[i] != -1 || excluded(i)) continue;
        for (int t = rowptr[i]; t < rowptr[i + 1]; ++t) {
            const int j = col[t];
            if (j != i && !excluded(j) && agg[j] >= 0) { tent[i] = agg[j]; break; }
        }
    }
    for (int i = 0; i < n; ++i)
        if (tent[i] >= 0) agg[i] = tent[i];
    // pass 3: whatever is left forms new aggregates with its free neighbours
    for (int i = 0; i < n; ++i) {
        if (agg[i] != -1 || excluded(i)) continue;
        agg[i] = na;
        for (int t = rowptr[i]; t < rowptr[i + 1]; ++t) {
            const int j = col[t];
            if (j != i && !excluded(j) && agg[j] == -1) agg[j] = na;
        }
        ++na;
    }
    *n_agg_out = na;
    return 0;
}
