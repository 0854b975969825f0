"""Host-compiled check of the curl-curl / rotational element routine (csrc/curlcurl_element.cuh, d = 2 and 3) against
oracle/curlcurl_oracle.py (itself checked against the sympy transcription of
src/solvers/stabilized_schur_pressurebc.py:85-160): residual and hand-derived Jacobian ≤1e-12 / 1e-11.
Groundwork for SURVEY §8(f) rank 4; test infrastructure only — the library never runs it on the CPU."""
import ctypes

import numpy as np
import pytest

from cfd_hemodynamic_b200.fem import quadrature as Q
from oracle import curlcurl_oracle as C
from oracle import ns_oracle as O
from oracle import simplex_oracle as S
from tests.test_simplex_host import _mesh, _p
from tests.test_tet_host import lib  # noqa: F401  (fixture: builds tests/host_simplex)


@pytest.mark.parametrize("d", [2, 3])
def test_curlcurl_routine_matches_oracle(lib, d):  # noqa: F811
    x, cells = _mesh(d)
    E, nv = cells.shape
    n = x.shape[0]
    h = S.cell_diameter(x, cells)
    rng = np.random.default_rng(17)
    u, p, un = rng.standard_normal((n, d)), rng.standard_normal(n), rng.standard_normal((n, d))
    un[0] = 0.0                                    # a vertex at rest: the tau_1 branch with |u_n| small near it
    f = np.array([0.1, -0.3, 0.2])
    par = dict(dt=0.02, rho=1.06, mu=0.035, eps0=O.EPS0)
    if d == 2:
        rules = [Q.triangle_gauss_jacobi(deg) for deg in (8, 7, 8, 7, 7, 6)]
    else:
        rules = [S.tet_gauss_jacobi(deg) for deg in (6, 5, 6, 5, 5, 4)]
    for b, (pts, wts) in enumerate(rules):
        pts, wts = np.ascontiguousarray(pts, dtype=np.float64), np.ascontiguousarray(wts, dtype=np.float64)
        assert lib.sxh_set_rule(d, b, _p(pts), _p(wts), len(wts)) == 0
    lib.sxh_set_params(par["dt"], par["rho"], par["mu"], _p(f), par["eps0"], 0.5, 1.0)
    nl = nv * nv
    Fe, Je = np.zeros(E * nl), np.zeros(E * nl * nl)
    sol = np.concatenate([u.reshape(-1), p])
    lib.cch_cells(d, E, n, _p(cells), _p(np.ascontiguousarray(x)), _p(h), _p(sol), _p(np.ascontiguousarray(un.reshape(-1))),
                  _p(Fe), _p(Je))
    Fe, Je = Fe.reshape(E, nl), Je.reshape(E, nl, nl)
    U, P, Un = u[cells], p[cells], un[cells]
    kw = dict(f=f[:d], **par)
    Fu, _ = C.element_F(x, cells, h, U, P, Un, rules[0], **kw)
    _, Fp = C.element_F(x, cells, h, U, P, Un, rules[1], **kw)
    F_ref = np.concatenate([Fu.reshape(E, -1), Fp], axis=1)
    assert np.abs(Fe - F_ref).max() < 1e-12 * np.abs(F_ref).max()
    J_ref = C.element_J(x, cells, h, U, P, Un, rules[2], rules[4], **kw)
    assert np.abs(Je - J_ref).max() < 1e-11 * np.abs(J_ref).max()


@pytest.mark.parametrize("d", [2, 3])
def test_curlcurl_facet_routine_matches_oracle(lib, d):  # noqa: F811
    """curlcurl_facet<D>: weak pressure + curl-form Nitsche terms (stabilized_schur_pressurebc.py:189-201) on
    every exterior facet of a small mesh; the Jacobian is checked against unit-vector differences of the
    (affine) oracle residual."""
    from types import SimpleNamespace
    from tests.test_tet_host import _coef8, _perturbed_cube
    from tests import common as T
    if d == 2:
        prob = T.make_problem(T.perturbed_square(3, 3, seed=5))
        x, cells = prob.x, prob.cells
        g3 = Q.interval_gauss(3)
        frule = (np.asarray(g3[0]).reshape(-1, 1), g3[1])
    else:
        x, cells = _perturbed_cube(2, seed=5)
        frule = S.triangle_facet_rule(4)
    n, nv = x.shape[0], d + 1
    h = S.cell_diameter(x, cells)
    pairs = S.exterior_facets(cells)
    m = len(pairs)
    rng = np.random.default_rng(19)
    u, p, un = rng.standard_normal((n, d)), rng.standard_normal(n), rng.standard_normal((n, d))
    coef = dict(pconst=3.7, a_n=1.0, beta_n=100.0)
    lib.sxh_set_params(0.02, 1.06, 0.035, _p(np.zeros(3)), O.EPS0, 0.5, 1.0)
    Fu, J = np.zeros((m, nv, d)), np.zeros((m, nv, nv, d, d))
    sol = np.concatenate([u.reshape(-1), p])
    c = lambda a, dt=np.float64: np.ascontiguousarray(a, dtype=dt)
    lib.cch_facets(d, m, n, _p(c(pairs, np.int32)), _p(c(cells, np.int32)), _p(c(x)), _p(c(h)), _p(sol), _p(c(un.reshape(-1))),
                   _p(_coef8(coef)), _p(c(frule[0])), _p(c(frule[1])), len(frule[1]), _p(Fu), _p(J))
    ns = SimpleNamespace(**coef)
    ce = pairs[:, 0]
    U, P, Un = u[cells][ce], p[cells][ce], un[cells][ce]
    ref = C.facet_F(x, cells, h, pairs, ns, U, P, Un, frule, 0.035)
    assert np.abs(Fu - ref).max() < 1e-12 * np.abs(ref).max()
    F0 = C.facet_F(x, cells, h, pairs, ns, np.zeros_like(U), P, Un, frule, 0.035)
    for b in range(nv):
        for l in range(d):
            E1 = np.zeros_like(U)
            E1[:, b, l] = 1.0
            dF = C.facet_F(x, cells, h, pairs, ns, E1, P, Un, frule, 0.035) - F0       # (m, a, k)
            assert np.abs(J[:, :, b, :, l] - dF).max() < 1e-12 * max(np.abs(dF).max(), 1.0)
