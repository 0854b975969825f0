"""Writes tests/golden/p1tet_small.npz: a small seeded case on P1 tetrahedra (mesh, fields,
parameters, cell and facet quadrature rules, one tagged facet set with every boundary term, two
velocity conditions sharing an edge and a pressure condition) with the oracle's Jacobian and residual
after the Dirichlet treatment (oracle/ns3d_oracle.py:assemble_system).

Like the 2-D files it pins the ORACLE (no reference outputs exist for this path);
the CPU oracle test, the host-compiled device code and the GPU parity test compare against it.

    python tests/golden/make_golden_tet.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ns3d_oracle as O3  # noqa: E402
from oracle import ns_oracle as O  # noqa: E402
from oracle import simplex_oracle as S  # noqa: E402

GOLDEN_TET = "p1tet_small.npz"
COEF = dict(a_p=1.0, a_g=1.0, pconst=0.7, a_n=1.0, beta_n=100.0, a_s=1.0, a_b=1.0, beta_b=0.2)
DEGREES = dict(Fu=6, Fp=5, uu=6, up=5, pu=5, pp=4)      # small rules keep the file small; the rules are stored in it


def build_case():
    x, cells = O3.unit_cube_tets(2)
    rng = np.random.default_rng(21)
    interior = (np.abs(x - 0.5) < 0.5 - 1e-12).all(axis=1)
    x = x.copy()
    x[interior] += 0.1 * (rng.random((int(interior.sum()), 3)) - 0.5)
    n = x.shape[0]
    rules = {k: S.tet_gauss_jacobi(d) for k, d in DEGREES.items()}
    pairs = S.exterior_facets(cells)
    fx = np.array([np.delete(x[cells[c]], lf, axis=0) for c, lf in pairs])
    fpairs = pairs[np.isclose(fx[:, :, 0], 1.0).all(axis=1) | np.isclose(fx[:, :, 2], 0.0).all(axis=1)]
    prob = O3.Problem3D(x=x, cells=cells, dt=0.02, rho=1.06, mu=0.035, f=np.array([0.1, -0.3, 0.2]), rules=rules,
                        facet_sets=[O.FacetSet(pairs=fpairs, **COEF)], facet_rule=S.triangle_facet_rule(4))
    u = np.stack([np.sin(2.1 * x[:, 0] + 0.3) * np.cos(1.7 * x[:, 1]), -np.cos(1.3 * x[:, 0]) * np.sin(2.3 * x[:, 2] + 0.2),
                  0.5 * np.sin(x[:, 1] + x[:, 2])], axis=1) + 0.01 * rng.standard_normal((n, 3))
    un = 0.9 * u + 0.05 * rng.standard_normal((n, 3))
    p = np.sin(1.1 * x[:, 0]) * x[:, 1] + 0.01 * rng.standard_normal(n)
    gu, gp = rng.standard_normal(3 * n), rng.standard_normal(n)
    n0 = np.nonzero(np.isclose(x[:, 0], 0.0))[0]
    n1 = np.nonzero(np.isclose(x[:, 1], 0.0))[0]
    n2 = np.nonzero(np.isclose(x[:, 0], 1.0))[0]
    bcs = [("u", n0, gu), ("u", n1, 2.0 * gu), ("p", n2, gp)]
    return prob, fpairs, bcs, u.reshape(-1), p, un.reshape(-1)


def bc_dof_lists(n, bcs):
    return [(3 * nodes[:, None] + np.arange(3)[None]).reshape(-1) if blk == "u" else 3 * n + nodes for blk, nodes, _ in bcs]


def bc_values(n, bcs):
    """g (4n): the last condition in the list wins on shared dofs."""
    g = np.zeros(4 * n)
    for (blk, nodes, vals), dofs in zip(bcs, bc_dof_lists(n, bcs)):
        g[dofs] = vals[dofs] if blk == "u" else vals[nodes]
    return g


def main():
    prob, fpairs, bcs, u, p, un = build_case()
    n = prob.n
    A, b = O3.assemble_system(prob, np.concatenate([u, p]), un, bc_values(n, bcs), bc_lists=bc_dof_lists(n, bcs))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), GOLDEN_TET)
    rules = {f"rule_{k}_{s}": v[i] for k, v in prob.rules.items() for i, s in enumerate(("pts", "wts"))}
    np.savez_compressed(out, x=prob.x, cells=prob.cells, u=u, p=p, un=un, indptr=A.indptr, indices=A.indices, data=A.data,
                        b=b, facet_pts=prob.facet_rule[0], facet_wts=prob.facet_rule[1], fpairs=fpairs, **rules)
    print("wrote", out, A.nnz, os.path.getsize(out))


if __name__ == "__main__":
    main()
