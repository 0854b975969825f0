"""CPU tests of the oracle itself (known-answer checks)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from cfd_hemodynamic_b200.fem import mesh as M
from cfd_hemodynamic_b200.fem import quadrature as Q
from oracle import ns_oracle as O
from tests import common as T

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "p1tri_small.npz")


def _small(with_facets=True):
    mesh = T.perturbed_square(4, 3, seed=2)
    prob = T.make_problem(mesh)
    if with_facets:
        ext = M.exterior_facet_indices(mesh.topology)
        pairs = mesh.topology.facet_cell_pairs(ext)
        prob.facet_sets = [O.FacetSet(pairs=pairs, a_p=1.0, a_g=1.0),
                           O.FacetSet(pairs=pairs[:5], pconst=3.0, a_s=2.0, a_n=2.0, beta_n=100.0, a_b=1.0, beta_b=0.2)]
    return mesh, prob


def test_jacobian_is_derivative_of_residual():
    """J = derivative(F) (stabilized_schur.py:185-187): F is polynomial in (u,p) for frozen
    u_n, so a complex-step derivative is exact to rounding."""
    mesh, prob = _small()
    rule = Q.triangle_gauss_jacobi(12)
    prob.rules = {k: rule for k in prob.rules}          # one rule: F and J blocks then share quadrature
    n = prob.n
    rng = np.random.default_rng(0)
    u, p, un = rng.standard_normal(2 * n), rng.standard_normal(n), rng.standard_normal(2 * n)
    A = O.assemble_J_raw(prob, u, p, un).toarray()
    eps = 1e-30
    Jcs = np.zeros_like(A)
    for j in range(3 * n):
        xx = np.concatenate([u, p]).astype(complex)
        xx[j] += 1j * eps
        Jcs[:, j] = O.assemble_F_raw(prob, xx[:2 * n], xx[2 * n:], un.astype(complex)).imag / eps
    assert np.linalg.norm(A - Jcs) <= 1e-14 * np.linalg.norm(Jcs)


def test_constant_pressure_nullspace_with_all_facet_term():
    """with the all-facet term of stabilized_schur.py:79 a constant pressure is in the
    kernel of the un-constrained Jacobian."""
    mesh, prob = _small(False)
    ext = M.exterior_facet_indices(mesh.topology)
    prob.facet_sets = [O.FacetSet(pairs=mesh.topology.facet_cell_pairs(ext), a_p=1.0, a_g=1.0)]
    n = prob.n
    u, p, un = T.smooth_fields(prob.x)
    A = O.assemble_J_raw(prob, u, p, un)
    v = np.zeros(3 * n)
    v[2 * n:] = 1.0
    assert np.linalg.norm(A @ v) <= 1e-13 * np.linalg.norm(A.data)


def test_rigid_state_has_zero_residual():
    """u = u_n = const, p = const, f = 0: every volume term vanishes (patch test)."""
    mesh, prob = _small(False)
    prob.f = np.zeros(2)
    n = prob.n
    u = np.tile([0.7, -0.2], n)
    b = O.assemble_F_raw(prob, u, np.full(n, 3.0), u.copy())
    # -p div(v) integrates to the boundary term p n.v: interior rows vanish
    x = prob.x
    interior = (np.abs(x - 0.5) < 0.5 - 1e-9).all(axis=1)
    rows = np.concatenate([np.repeat(interior, 2), interior])
    assert np.abs(b[rows]).max() < 1e-13


def test_dirichlet_semantics():
    """Rows/cols zeroed, diagonal = number of DirichletBC objects holding the dof,
    residual rows = x - g, lifting only while x violates the BC."""
    mesh, prob = _small()
    n = prob.n
    x = prob.x
    left = np.nonzero(np.isclose(x[:, 0], 0.0))[0]
    bottom = np.nonzero(np.isclose(x[:, 1], 0.0))[0]
    rng = np.random.default_rng(1)
    g0, g1 = rng.standard_normal(2 * n), rng.standard_normal(2 * n)
    prob.bcs = T.oracle_bcs(prob, [("u", left, g0), ("u", bottom, g1)])
    u, p, un = T.smooth_fields(x)
    A = O.assemble_J(prob, u, p, un)
    marker, g, mult = O.bc_arrays(prob)
    assert mult.max() == 2.0                     # the corner node is in both BCs
    d = A.diagonal()
    assert np.array_equal(d[marker], mult[marker])
    Ad = A.toarray()
    off = Ad - np.diag(d)
    assert np.abs(off[marker]).max() == 0.0 and np.abs(off[:, marker]).max() == 0.0
    corner = np.intersect1d(left, bottom)[0]
    assert g[2 * corner] == g1[2 * corner]       # last BC in the list wins
    xx = np.concatenate([u, p])
    b = O.assemble_F(prob, xx, un)
    assert np.allclose(b[marker], xx[marker] - g[marker], rtol=0, atol=0)
    xx[marker] = g[marker]
    b2 = O.assemble_F(prob, xx, un)
    raw = O.assemble_F_raw(prob, xx[:2 * n], xx[2 * n:], un)
    assert np.array_equal(b2[~marker], raw[~marker])   # no lifting once x satisfies the BC


def test_newton_poiseuille_convergence():
    """Known answer (reference src/scenarios/unit_square.py:100-105): channel flow
    u = (4 y (1 - y), 0), p = -8 mu x.  One huge time step from the exact state lands on the
    discrete steady state; its error must fall at second order under refinement."""
    errs, slopes = [], []
    for nx in (8, 16):
        mesh = M.create_unit_square(None, nx, nx)
        prob = T.make_problem(mesh, dt=1e3, rho=1.0, mu=1.0, f=(0.0, 0.0))
        x = prob.x
        n = prob.n
        ext = M.exterior_facet_indices(mesh.topology)
        prob.facet_sets = [O.FacetSet(pairs=mesh.topology.facet_cell_pairs(ext), a_p=1.0, a_g=1.0)]
        bnd = np.nonzero(np.isclose(x[:, 0], 0) | np.isclose(x[:, 0], 1) | np.isclose(x[:, 1], 0) | np.isclose(x[:, 1], 1))[0]
        ue = np.zeros(2 * n)
        ue[0::2] = 4 * x[:, 1] * (1 - x[:, 1])
        prob.bcs = T.oracle_bcs(prob, [("u", bnd, ue)])
        xk, its, reason = O.newton_solve(prob, np.zeros(3 * n), ue.copy(), rtol=1e-10)
        assert reason > 0 and its <= 6
        errs.append(np.linalg.norm(xk[:2 * n] - ue) / np.linalg.norm(ue))
        p = xk[2 * n:].reshape(nx + 1, nx + 1)
        slopes.append(np.polyfit(np.linspace(0, 1, nx + 1), p[nx // 2], 1)[0])
    assert errs[1] < errs[0] / 2.5 and errs[1] < 0.05, errs
    assert abs(slopes[1] + 8.0) < abs(slopes[0] + 8.0) and abs(slopes[1] + 8.0) < 0.8, slopes


def test_outlet_flux_linear_field():
    mesh = M.create_unit_square(None, 5, 7)
    prob = T.make_problem(mesh)
    outlet = M.locate_entities_boundary(mesh, 1, lambda X: np.isclose(X[0], 1.0))
    un = np.zeros(2 * prob.n)
    un[0::2] = 2.0 + prob.x[:, 1]
    q = O.outlet_flux(prob, mesh.topology.facet_cell_pairs(outlet), un)
    assert abs(q - 2.5) < 1e-14


@pytest.mark.parametrize("cell_type", ["triangle", "quadrilateral"])
def test_golden_vectors(cell_type):
    """The committed golden cases (tests/golden/make_golden.py; P1 triangles and Q1
    quadrilaterals) are reproduced bit-for-bit in pattern and to 1e-13 in values by the
    current oracle."""
    from tests.golden.make_golden import GOLDEN_FILES, build_case
    gold = np.load(os.path.join(os.path.dirname(GOLDEN), GOLDEN_FILES[cell_type]))
    mesh, prob, fsets, bcs, u, p, un = build_case(cell_type)
    for k in prob.rules:                       # use the committed rules: golden pins arithmetic, not tables
        prob.rules[k] = (gold[f"rule_{k}_pts"], gold[f"rule_{k}_wts"])
    assert np.array_equal(gold["cells"], prob.cells) and np.array_equal(gold["x"], prob.x)
    A = O.assemble_J(prob, gold["u"], gold["p"], gold["un"])
    assert np.array_equal(A.indptr, gold["indptr"]) and np.array_equal(A.indices, gold["indices"])
    assert np.linalg.norm(A.data - gold["data"]) <= 1e-13 * np.linalg.norm(gold["data"])
    b = O.assemble_F(prob, np.concatenate([gold["u"], gold["p"]]), gold["un"])
    assert np.linalg.norm(b - gold["b"]) <= 1e-13 * np.linalg.norm(gold["b"])


def test_c_restatement_matches_numpy_oracle():
    """oracle/c/p1tri_cells.c (OpenMP, FFCx-style loops) vs the numpy restatement."""
    from oracle import c_oracle
    mesh, prob = _small()
    n = prob.n
    u, p, un = T.smooth_fields(prob.x)
    Ae, Fe = c_oracle.element_tensors(prob, u, p, un)
    Ae_ref = O.element_matrices(prob, u, p, un)
    U, P, Un = O._gather(prob, u, p, un)
    Fu, _ = O.element_F(prob, U, P, Un, prob.rules["Fu"])
    _, Fp = O.element_F(prob, U, P, Un, prob.rules["Fp"])
    Fe_ref = np.concatenate([Fu.reshape(-1, 6), Fp], axis=1)
    assert np.linalg.norm(Ae - Ae_ref) <= 1e-13 * np.linalg.norm(Ae_ref)
    assert np.linalg.norm(Fe - Fe_ref) <= 1e-13 * np.linalg.norm(Fe_ref)
    x = prob.x
    left = np.nonzero(np.isclose(x[:, 0], 0.0))[0]
    prob.bcs = T.oracle_bcs(prob, [("u", left, np.random.default_rng(3).standard_normal(2 * n))])
    fa = c_oracle.FastAssembler(prob)
    xx = np.concatenate([u, p])
    A = fa.J(u, p, un)
    A_ref = O.assemble_J(prob, u, p, un)
    assert abs(A - A_ref).max() <= 1e-12 * abs(A_ref).max()
    assert np.linalg.norm(fa.F(xx, un) - O.assemble_F(prob, xx, un)) <= 1e-12 * np.linalg.norm(O.assemble_F(prob, xx, un))
