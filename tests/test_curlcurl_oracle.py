"""Groundwork for SURVEY §8(f) rank 4: the curl-curl / rotational formulation of
src/solvers/stabilized_schur_pressurebc.py:85-160 restated in oracle/curlcurl_oracle.py, checked against the literal
sympy transcription of the form text (triangles and tetrahedra) and by known answers."""
import numpy as np
import pytest

from cfd_hemodynamic_b200.fem import quadrature as Q
from oracle import curlcurl_oracle as C
from oracle import ns_oracle as O
from oracle import simplex_oracle as S
from oracle.form_mirror import CurlCurlForms

PAR = dict(dt=0.02, rho=1.06, mu=0.035, eps0=O.EPS0)


def _cell(d):
    if d == 2:
        return np.array([[0.0, 0.1], [1.0, 0.0], [0.2, 0.9]]), (Q.triangle_gauss_jacobi(6), Q.triangle_gauss_jacobi(5))
    X = np.array([[0.0, 0.1, 0.0], [1.0, 0.0, 0.1], [0.2, 0.9, 0.0], [0.1, 0.2, 0.8]])
    return X, (S.tet_gauss_jacobi(6), S.tet_gauss_jacobi(5))


@pytest.mark.parametrize("d", [2, 3])
def test_curlcurl_cell_residual_matches_form_text(d):
    X, (rule_u, rule_p) = _cell(d)
    cells = np.arange(d + 1, dtype=np.int32)[None, :]
    h = S.cell_diameter(X, cells)
    rng = np.random.default_rng(3 + d)
    U, P, Un = rng.standard_normal((d + 1, d)), rng.standard_normal(d + 1), rng.standard_normal((d + 1, d))
    f = (0.1, -0.3, 0.2)[:d]
    cf = CurlCurlForms(X, Un, float(h[0]), f=f, **PAR)
    Fu, _ = C.element_F(X, cells, h, U[None], P[None], Un[None], rule_u, f=f, **PAR)
    _, Fp = C.element_F(X, cells, h, U[None], P[None], Un[None], rule_p, f=f, **PAR)
    Fu_m, Fp_m = cf.cell_residual(U, P, rule_u, rule_p)
    scale = max(np.abs(Fu).max(), np.abs(Fp).max())
    assert np.abs(Fu[0] - Fu_m).max() < 1e-12 * scale and np.abs(Fp[0] - Fp_m).max() < 1e-12 * scale
    # complex-step Jacobian of the oracle vs complex step of the mirrored residual
    J = C.element_J(X, cells, h, U[None], P[None], Un[None], rule_u, rule_p, f=f, **PAR)[0]
    nl = (d + 1) * (d + 1)
    x0 = np.concatenate([U.reshape(-1), P])
    Jm = np.zeros((nl, nl))
    for j in range(nl):
        xc = x0.astype(complex)
        xc[j] += 1e-30j
        fu, fp = cf.cell_residual(xc[:d * (d + 1)].reshape(d + 1, d), xc[d * (d + 1):], rule_u, rule_p)
        Jm[:d * (d + 1), j] = fu.reshape(-1).imag / 1e-30
        Jm[d * (d + 1):, j] = fp.imag / 1e-30
    assert np.abs(J - Jm).max() < 1e-11 * np.abs(J).max()


@pytest.mark.parametrize("d", [2, 3])
def test_constant_and_hydrostatic_known_answers(d):
    """(1) Constant velocity c = u = u_n, constant pressure p0, f = 0: every volume term vanishes except
    -(p + rho |u_m|^2 / 2) div v (the rotational form carries the Bernoulli pressure), so F_u[a] =
    -(p0 + rho |c|^2 / 2) |K| grad(phi_a) and F_p = 0.
    (2) Hydrostatic state u = u_n = 0, p = rho f . x: the strong residual R vanishes, F_p = 0 and the velocity
    rows sum to -rho f |K| over the cell's vertices (partition of unity)."""
    X, (rule_u, rule_p) = _cell(d)
    cells = np.arange(d + 1, dtype=np.int32)[None, :]
    h = S.cell_diameter(X, cells)
    c = np.array([0.7, -0.4, 0.3])[:d]
    U = np.tile(c, (d + 1, 1))
    P = np.full(d + 1, 2.5)
    Fu, Fp = C.element_F(X, cells, h, U[None], P[None], U[None], rule_u, f=np.zeros(d), **PAR)
    # constant fields: every volume term vanishes except -(p + rho |c|^2/2) div v, which integrates to
    # -(p + rho |c|^2 / 2) |K| grad(phi_a): sums to zero over the cell's vertices (partition of unity)
    assert np.abs(Fp).max() < 1e-14
    assert np.abs(Fu[0].sum(axis=0)).max() < 1e-14
    _, dphi = S.simplex_geometry(X, cells)
    vol = np.abs(np.linalg.det((X[1:] - X[0]).T)) / (2.0 if d == 2 else 6.0)
    assert np.abs(Fu[0] + (2.5 + PAR["rho"] * 0.5 * c @ c) * vol * dphi[0]).max() < 1e-13
    # hydrostatic: u = u_n = 0, p = rho f . x
    fvec = np.array([0.1, -0.3, 0.2])[:d]
    P = PAR["rho"] * X @ fvec
    Z = np.zeros((d + 1, d))
    Fu, Fp = C.element_F(X, cells, h, Z[None], P[None], Z[None], rule_u, f=fvec, **PAR)
    assert np.abs(Fp).max() < 1e-14                                           # R = 0, div u = 0
    # F_u[a] = -int p grad(phi_a) - rho f int phi_a ; summed over a: -rho f |K| (grad of the partition of unity is 0)
    assert np.abs(Fu[0].sum(axis=0) + PAR["rho"] * fvec * vol).max() < 1e-14


@pytest.mark.parametrize("d", [2, 3])
def test_curlcurl_facet_terms_match_form_text(d):
    """Weak pressure + Nitsche terms of stabilized_schur_pressurebc.setup (:189-201) on every local facet."""
    from types import SimpleNamespace
    X, _ = _cell(d)
    cells = np.arange(d + 1, dtype=np.int32)[None, :]
    h = S.cell_diameter(X, cells)
    rng = np.random.default_rng(11 + d)
    U, P, Un = rng.standard_normal((d + 1, d)), rng.standard_normal(d + 1), rng.standard_normal((d + 1, d))
    cf = CurlCurlForms(X, Un, float(h[0]), f=(0.0,) * d, **PAR)
    frule = (np.asarray(Q.interval_gauss(3)[0]).reshape(-1, 1), Q.interval_gauss(3)[1]) if d == 2 else S.triangle_facet_rule(4)
    coef = SimpleNamespace(pconst=3.7, a_n=1.0, beta_n=100.0)
    for lf in range(d + 1):
        pairs = np.array([[0, lf]], dtype=np.int32)
        nrm, scale = S.facet_geometry(X, cells, pairs)
        got = C.facet_F(X, cells, h, pairs, coef, U[None], P[None], Un[None], frule, PAR["mu"])[0]
        ref = cf.facet_residual(U, P, lf, nrm[0], float(scale[0]), frule, pconst=3.7, a_n=1.0, beta_n=100.0)
        assert np.abs(got - ref).max() < 1e-12 * np.abs(ref).max()
