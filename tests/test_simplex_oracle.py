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
