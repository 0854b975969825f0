"""Host-compiled check of the dimension-generic P1 simplex element routines
(csrc/simplex_element.cuh: moment factorisation, D = 2 and D = 3) against the FFCx-style oracle
(oracle/simplex_oracle.py, itself checked against the sympy transcription of the form).
Groundwork for the tetrahedral kernels; tolerance 1e-12 relative per element tensor."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from cfd_hemodynamic_b200.fem import quadrature as Q
from oracle import ns_oracle as O
from oracle import simplex_oracle as S

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(HERE, "host_simplex", "simplex_host.cpp")
    out_dir = os.path.join(HERE, "host_simplex", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libsimplexhost.so")
    deps = [src] + [os.path.join(HERE, "..", "cfd_hemodynamic_b200", "csrc", f) for f in ("simplex_element.cuh", "tet_items.cuh", "curlcurl_element.cuh", "hemo_rules.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src], check=True)
    L = ctypes.CDLL(so)
    L.sxh_set_params.argtypes = [ctypes.c_double] * 3 + [ctypes.c_void_p] + [ctypes.c_double] * 3
    return L


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _mesh(d, seed=0):
    """A few randomly perturbed simplices sharing vertices (no global consistency needed here)."""
    rng = np.random.default_rng(seed)
    if d == 2:
        x = np.array([[0, 0], [1, 0], [0, 1], [1, 1], [0.5, 1.7]], dtype=float)
        cells = np.array([[0, 1, 2], [1, 3, 2], [2, 3, 4]], dtype=np.int32)
    else:
        x = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1], [0.3, 0.2, 1.8]], dtype=float)
        cells = np.array([[0, 1, 2, 3], [1, 2, 3, 4], [3, 2, 4, 5]], dtype=np.int32)
    x = x + 0.1 * rng.standard_normal(x.shape)
    return x, cells


@pytest.mark.parametrize("theta,a0", [(0.5, 1.0), (1.0, 1.5)])
@pytest.mark.parametrize("d", [2, 3])
def test_simplex_routines_match_oracle(lib, d, theta, a0):
    x, cells = _mesh(d)
    E, nv = cells.shape
    n = x.shape[0]
    h = S.cell_diameter(x, cells)
    rng = np.random.default_rng(7)
    u, p, un = rng.standard_normal((n, d)), rng.standard_normal(n), rng.standard_normal((n, d))
    uh = 2.0 * un - 0.5 * rng.standard_normal((n, d)) if theta == 1.0 else un.copy()
    f = np.array([0.1, -0.3, 0.2])
    par = dict(dt=0.02, rho=1.06, mu=0.035, f=f[:d], eps0=O.EPS0, theta=theta, a0=a0)
    if d == 2:
        rules = [Q.triangle_rule(deg) for deg in (12, 11, 12, 11, 11, 10)]
    else:
        rules = [S.tet_gauss_jacobi(deg) for deg in (12, 11, 12, 11, 11, 10)]       # 7^3 / 6^3 points
    for b, (pts, wts) in enumerate(rules):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        wts = np.ascontiguousarray(wts, dtype=np.float64)
        assert lib.sxh_set_rule(d, b, _p(pts), _p(wts), len(wts)) == 0
    lib.sxh_set_params(par["dt"], par["rho"], par["mu"], _p(f), par["eps0"], theta, a0)
    B = d + 1
    Ae = np.zeros(E * nv * nv * B * B)
    Fe = np.zeros(E * nv * B)
    sol = np.concatenate([u.reshape(-1), p])
    lib.sxh_cells(d, E, n, _p(cells), _p(np.ascontiguousarray(x)), _p(h), _p(sol), _p(np.ascontiguousarray(un.reshape(-1))),
                  _p(np.ascontiguousarray(uh.reshape(-1))), _p(Ae), _p(Fe))
    Ae = Ae.reshape(E, nv, nv, B, B)
    Fe = Fe.reshape(E, nv, B)
    U, P, Un, Uh = u[cells], p[cells], un[cells], uh[cells]
    kw = dict(Uh=Uh, **par)
    Fu, _ = S.element_F(x, cells, h, U, P, Un, rules[0], **kw)
    _, Fp = S.element_F(x, cells, h, U, P, Un, rules[1], **kw)
    Juu, _, _, _ = S.element_J(x, cells, h, U, P, Un, rules[2], **kw)
    _, Jup, _, _ = S.element_J(x, cells, h, U, P, Un, rules[3], **kw)
    _, _, Jpu, _ = S.element_J(x, cells, h, U, P, Un, rules[4], **kw)
    _, _, _, Jpp = S.element_J(x, cells, h, U, P, Un, rules[5], **kw)
    tol = 1e-12
    assert np.abs(Fe[:, :, :d] - Fu).max() < tol * np.abs(Fu).max()
    assert np.abs(Fe[:, :, d] - Fp).max() < tol * max(np.abs(Fp).max(), np.abs(Fu).max())
    # Ae[e, a, b, ri, ci]  vs  Juu[e, a, k, b, l] etc.
    assert np.abs(Ae[:, :, :, :d, :d] - Juu.transpose(0, 1, 3, 2, 4)).max() < tol * np.abs(Juu).max()
    assert np.abs(Ae[:, :, :, :d, d] - Jup.transpose(0, 1, 3, 2)).max() < tol * np.abs(Jup).max()
    assert np.abs(Ae[:, :, :, d, :d] - Jpu).max() < tol * np.abs(Jpu).max()
    assert np.abs(Ae[:, :, :, d, d] - Jpp).max() < tol * np.abs(Jpp).max()
