"""Writes tests/golden/p1tri_small.npz and q1quad_small.npz: small seeded cases (mesh,
fields, parameters, quadrature rules) with the oracle's Jacobian values and residual on
P1 triangles and on Q1 quadrilaterals.

No reference outputs exist for this path (the reference ships no tests or
fixtures and DOLFINx/PETSc cannot be imported here — SURVEY.md §8(c)), so the
golden vectors pin the ORACLE (regression) rather than the reference; both the
CPU oracle tests and the GPU parity tests compare against them.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from cfd_hemodynamic_b200.fem import mesh as M  # noqa: E402
from oracle import ns_oracle as O  # noqa: E402
from tests import common as T  # noqa: E402


def build_case(cell_type="triangle"):
    mesh = T.perturbed_square(5, 4, seed=11, cell_type=cell_type)
    prob = T.make_problem(mesh, dt=0.02, rho=1.06, mu=0.035, f=(0.1, -0.3))
    ext = M.exterior_facet_indices(mesh.topology)
    inlet = M.locate_entities_boundary(mesh, 1, lambda X: np.isclose(X[0], 0.0))
    outlet = M.locate_entities_boundary(mesh, 1, lambda X: np.isclose(X[0], 1.0))
    fsets = [(ext, dict(a_p=1.0, a_g=1.0)),
             (inlet, dict(pconst=5.0, a_n=1.0, beta_n=100.0)),
             (outlet, dict(pconst=0.7, a_s=1.0, a_b=1.0, beta_b=0.2))]
    x = prob.x
    n = prob.n
    walls = np.nonzero(np.isclose(x[:, 1], 0.0) | np.isclose(x[:, 1], 1.0))[0]
    rng = np.random.default_rng(7)
    bcs = [("u", walls, rng.standard_normal(2 * n))]
    prob.facet_sets = [O.FacetSet(pairs=mesh.topology.facet_cell_pairs(f), **c) for f, c in fsets]
    prob.bcs = T.oracle_bcs(prob, bcs)
    u, p, un = T.smooth_fields(x, seed=9)
    return mesh, prob, fsets, bcs, u, p, un


GOLDEN_FILES = {"triangle": "p1tri_small.npz", "quadrilateral": "q1quad_small.npz"}


def main(which=("triangle", "quadrilateral")):
    for cell_type in which:
        mesh, prob, fsets, bcs, u, p, un = build_case(cell_type)
        A = O.assemble_J(prob, u, p, un)
        b = O.assemble_F(prob, np.concatenate([u, p]), un)
        out = os.path.join(os.path.dirname(os.path.abspath(__file__)), GOLDEN_FILES[cell_type])
        rules = {f"rule_{k}_{s}": v[i] for k, v in prob.rules.items() for i, s in enumerate(("pts", "wts"))}
        np.savez_compressed(out, x=prob.x, cells=prob.cells, u=u, p=p, un=un, indptr=A.indptr, indices=A.indices,
                            data=A.data, b=b, facet_pts=prob.facet_rule[0], facet_wts=prob.facet_rule[1], **rules)
        print("wrote", out, A.nnz)


if __name__ == "__main__":
    # the committed triangle file is only rewritten on request: python make_golden.py triangle
    main(tuple(sys.argv[1:]) or ("quadrilateral",))
