"""Dimension-generic simplex oracle (oracle/simplex_oracle.py): identical to the triangle oracle for
d = 2, and checked against the literal sympy transcription of the form and by complex-step
differentiation on tetrahedra (d = 3) — groundwork for the tetrahedral kernels."""
import numpy as np
import pytest

from cfd_hemodynamic_b200.fem import quadrature as Q
from oracle import ns_oracle as O
from oracle import simplex_oracle as S
from oracle.form_mirror import SimplexForms
from tests import common as T

PAR = dict(dt=0.02, rho=1.06, mu=0.035, f=(0.1, -0.3), eps0=O.EPS0)


def test_generic_code_reproduces_triangle_oracle():
    mesh = T.perturbed_square(4, 3, seed=2)
    prob = T.make_problem(mesh, dt=PAR["dt"], rho=PAR["rho"], mu=PAR["mu"], f=PAR["f"])
    u, p, un = T.smooth_fields(prob.x)
    U, P, Un = O._gather(prob, u, p, un)
    for theta, a0 in ((0.5, 1.0), (1.0, 1.5)):
        prob.theta, prob.a0 = theta, a0
        rule = prob.rules["uu"]
        ref = O.element_F(prob, U, P, Un, rule) + O.element_J(prob, U, P, Un, rule)
        got = (S.element_F(prob.x, prob.cells, prob.h, U, P, Un, rule, theta=theta, a0=a0, **PAR)
               + S.element_J(prob.x, prob.cells, prob.h, U, P, Un, rule, theta=theta, a0=a0, **PAR))
        for r, g in zip(ref, got):
            assert np.abs(r - g).max() <= 1e-14 * np.abs(r).max()


def test_tet_rule_integrates_polynomials():
    pts, wts = S.tet_gauss_jacobi(6)
    assert abs(wts.sum() - 1.0 / 6.0) < 1e-15
    # int x^a y^b z^c over the reference tetrahedron = a! b! c! / (a+b+c+3)!
    from math import factorial as fa
    for a, b, c in ((1, 0, 0), (2, 1, 0), (1, 1, 1), (2, 2, 2), (0, 3, 3)):
        exact = fa(a) * fa(b) * fa(c) / fa(a + b + c + 3)
        assert abs(np.sum(wts * pts[:, 0] ** a * pts[:, 1] ** b * pts[:, 2] ** c) - exact) < 1e-15


@pytest.mark.parametrize("theta,a0", [(0.5, 1.0), (1.0, 1.5)])
def test_tetrahedron_matches_form_text(theta, a0):
    X = np.array([[0.0, 0.1, 0.0], [1.0, 0.0, 0.1], [0.2, 0.9, 0.0], [0.1, 0.2, 0.8]])
    cells = np.arange(4, dtype=np.int32)[None, :]
    h = S.cell_diameter(X, cells)
    rng = np.random.default_rng(4)
    U, P, Un = rng.standard_normal((4, 3)), rng.standard_normal(4), rng.standard_normal((4, 3))
    Uh = 2.0 * Un - 0.5 * rng.standard_normal((4, 3)) if theta == 1.0 else None
    par = dict(dt=0.02, rho=1.06, mu=0.035, f=(0.1, -0.3, 0.2), eps0=O.EPS0, theta=theta, a0=a0)
    rule_u, rule_p = S.tet_gauss_jacobi(6), S.tet_gauss_jacobi(5)
    cf = SimplexForms(X, Un, float(h[0]), Uh=Uh, **par)
    kw = dict(Uh=None if Uh is None else Uh[None], **par)
    Fu, _ = S.element_F(X, cells, h, U[None], P[None], Un[None], rule_u, **kw)
    _, Fp = S.element_F(X, cells, h, U[None], P[None], Un[None], rule_p, **kw)
    Fu_m, Fp_m = cf.cell_residual(U, P, rule_u, rule_p)
    scale = max(np.abs(Fu).max(), np.abs(Fp).max())
    assert np.abs(Fu[0] - Fu_m).max() < 1e-12 * scale and np.abs(Fp[0] - Fp_m).max() < 1e-12 * scale
    # hand-derived 16 x 16 Jacobian vs complex step of the mirrored residual (one rule for every block)
    Juu, Jup, Jpu, Jpp = S.element_J(X, cells, h, U[None], P[None], Un[None], rule_u, **kw)
    A = np.zeros((16, 16))
    A[:12, :12] = Juu[0].reshape(12, 12)
    A[:12, 12:] = Jup[0].reshape(12, 4)
    A[12:, :12] = Jpu[0].reshape(4, 12)
    A[12:, 12:] = Jpp[0]
    x0 = np.concatenate([U.reshape(-1), P])
    J = np.zeros((16, 16))
    for j in range(16):
        xc = x0.astype(complex)
        xc[j] += 1e-30j
        fu, fp = cf.cell_residual(xc[:12].reshape(4, 3), xc[12:], rule_u, rule_u)
        J[:12, j] = fu.reshape(-1).imag / 1e-30
        J[12:, j] = fp.imag / 1e-30
    assert np.abs(A - J).max() < 1e-11 * np.abs(A).max()


# ---- exterior-facet integrals ------------------------------------------------------------------
FACET_COEFS = [
    dict(a_p=1.0, a_g=1.0),                                              # stabilized_schur.py:79
    dict(pconst=3.7, a_n=1.0, beta_n=100.0),                             # pressure_backflow.py:192-201 (inlet)
    dict(pconst=-1.2, a_s=1.0, a_b=1.0, beta_b=0.2),                     # :208-217 (outlet)
    dict(a_p=2.0, a_g=2.0, pconst=1.0, a_n=2.0, beta_n=10.0, a_s=2.0, a_b=2.0, beta_b=0.5),   # double setup()
]


@pytest.mark.parametrize("coef", FACET_COEFS)
def test_generic_facet_code_reproduces_triangle_oracle(coef):
    mesh = T.perturbed_square(4, 3, seed=2)
    prob = T.make_problem(mesh, dt=PAR["dt"], rho=PAR["rho"], mu=PAR["mu"], f=PAR["f"])
    pairs = S.exterior_facets(prob.cells)
    assert len(pairs) == 2 * (4 + 3)
    fs = O.FacetSet(pairs=pairs, **coef)
    u, p, un = T.smooth_fields(prob.x)
    U, P, Un = O._gather(prob, u, p, un)
    ce = pairs[:, 0]
    for theta in (0.5, 1.0):
        prob.theta = theta
        ref = O.facet_F(prob, fs, U[ce], P[ce], Un[ce])
        got = S.facet_F(prob.x, prob.cells, prob.h, pairs, fs, U[ce], P[ce], Un[ce], prob.facet_rule,
                        PAR["rho"], PAR["mu"], theta)
        assert np.abs(ref - got).max() <= 1e-14 * np.abs(ref).max()
    assert abs(S.outlet_flux(prob.x, prob.cells, pairs, un) - O.outlet_flux(prob, pairs, un)) < 1e-14


@pytest.mark.parametrize("coef", FACET_COEFS)
@pytest.mark.parametrize("lf", [0, 1, 2, 3])
def test_tetrahedron_facets_match_form_text(coef, lf):
    X = np.array([[0.0, 0.1, 0.0], [1.0, 0.0, 0.1], [0.2, 0.9, 0.0], [0.1, 0.2, 0.8]])
    cells = np.arange(4, dtype=np.int32)[None, :]
    h = S.cell_diameter(X, cells)
    rng = np.random.default_rng(5 + lf)
    U, P, Un = rng.standard_normal((4, 3)), rng.standard_normal(4), rng.standard_normal((4, 3))
    par = dict(dt=0.02, rho=1.06, mu=0.035, f=(0.1, -0.3, 0.2), eps0=O.EPS0, theta=0.5, a0=1.0)
    cf = SimplexForms(X, Un, float(h[0]), **par)
    pairs = np.array([[0, lf]], dtype=np.int32)
    nrm, scale = S.facet_geometry(X, cells, pairs)
    # the normal is orthogonal to the facet, of unit length and points away from the opposite vertex
    fv = S.facet_vertices(3)[lf]
    assert abs(np.linalg.norm(nrm[0]) - 1.0) < 1e-15
    assert np.abs((X[fv[1:]] - X[fv[0]]) @ nrm[0]).max() < 1e-15
    assert (X[fv[0]] - X[lf]) @ nrm[0] > 0.0
    rule = S.triangle_facet_rule(4)
    fs = O.FacetSet(pairs=pairs, **coef)
    got = S.facet_F(X, cells, h, pairs, fs, U[None], P[None], Un[None], rule, par["rho"], par["mu"], 0.5)[0]
    ref = cf.facet_residual(U, P, lf, nrm[0], float(scale[0]), rule, **coef)
    assert np.abs(got - ref).max() < 1e-12 * np.abs(ref).max()


def test_triangle_facet_rules_and_closed_surface():
    from math import factorial as fa
    for deg in (2, 4):
        pts, wts = S.triangle_facet_rule(deg)
        for a in range(deg + 1):
            for b in range(deg + 1 - a):
                assert abs(np.sum(wts * pts[:, 0] ** a * pts[:, 1] ** b) - fa(a) * fa(b) / fa(a + b + 2)) < 1e-14
    # a constant velocity has no net flux through the closed surface of a tetrahedral mesh
    from oracle.ns3d_oracle import unit_cube_tets
    x, cells = unit_cube_tets(2)
    pairs = S.exterior_facets(cells)
    assert len(pairs) == 6 * 2 * 2 * 2
    un = np.tile([0.3, -1.1, 0.7], x.shape[0])
    assert abs(S.outlet_flux(x, cells, pairs, un)) < 1e-14
    top = pairs[[np.allclose(x[np.delete(cells[c], lf)][:, 2], 1.0) for c, lf in pairs]]
    assert abs(S.outlet_flux(x, cells, top, un) - 0.7) < 1e-14
