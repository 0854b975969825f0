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
