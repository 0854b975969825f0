"""P2-P2 triangle path, CPU checks (no GPU): the host+device element routines of csrc/p2_element.cuh — the arithmetic the CUDA
kernels of assembly_p2.cu wrap — compiled with g++ (tests/host_p2) and compared with oracle/pk_oracle.py through the same SoA
element buffers as the device.  Tolerance 1e-12 relative Frobenius (north_star)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from cfd_hemodynamic_b200.fem import discretization as D
from cfd_hemodynamic_b200.fem import mesh as M
from cfd_hemodynamic_b200.fem import quadrature as Q
from oracle import ns_oracle as O
from oracle import pk_oracle as PK
from tests import common as T

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-12


class _Params(ctypes.Structure):
    _fields_ = [("dt", ctypes.c_double), ("rho", ctypes.c_double), ("mu", ctypes.c_double), ("f", ctypes.c_double * 2),
                ("eps0", ctypes.c_double)]


class _Coef(ctypes.Structure):
    _fields_ = [(k, ctypes.c_double) for k in ("a_p", "pconst", "a_g", "a_s", "a_n", "beta_n", "a_b", "beta_b")]


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(HERE, "host_p2", "p2_host.cpp")
    out_dir = os.path.join(HERE, "host_p2", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libp2host.so")
    deps = [src] + [os.path.join(HERE, "..", "cfd_hemodynamic_b200", "csrc", f) for f in ("p2_element.cuh", "hemo_rules.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src], check=True)
    L = ctypes.CDLL(so)
    L.p2h_flux.restype = ctypes.c_double
    L.p2h_set_time_scheme.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
    return L


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


P2_DEGREE = {"Fu": 20, "Fp": 18, "uu": 20, "up": 18, "pu": 18, "pp": 16}


def _problem(nx=4, ny=3, seed=0, degrees=P2_DEGREE, **pk):
    mesh = T.perturbed_square(nx, ny, seed=seed)
    x, cells6 = D.p2_nodes(mesh)
    rules = {k: Q.triangle_rule(d) for k, d in degrees.items()}
    kw = dict(dt=0.01, rho=1.3, mu=0.02, f=np.array([0.3, -0.2]))
    kw.update(pk)
    prob = O.Problem(x=x, cells=cells6, h=PK.cell_diameter(x, cells6), rules=rules, facet_rule=Q.interval_gauss(4), **kw)
    return mesh, prob


def _load(lib, prob):
    for k, bid in T.BLOCK_ID.items():
        pts, wts = (np.ascontiguousarray(a, dtype=np.float64) for a in prob.rules[k])
        lib.p2h_set_rule(bid, _p(pts), _p(wts), len(wts))
    s, w = (np.ascontiguousarray(a, dtype=np.float64) for a in prob.facet_rule)
    lib.p2h_set_facet_rule(_p(s), _p(w), len(w))
    par = _Params(prob.dt, prob.rho, prob.mu, (ctypes.c_double * 2)(*prob.f), prob.eps0)
    lib.p2h_set_params(ctypes.byref(par))
    uh = None if prob.uh is None else np.ascontiguousarray(prob.uh, dtype=np.float64)
    prob._keep = uh
    lib.p2h_set_time_scheme(prob.theta, prob.a0, None if uh is None else _p(uh))


def _soa_to_ae(Ae, E):
    """SoA [(a*6+b)*9 + ri*3+ci][E] -> (E, 18, 18) in the oracle's local order [u(a,k) -> 2a+k | p(a) -> 12+a]."""
    A = Ae.reshape(6, 6, 3, 3, E)
    out = np.zeros((E, 18, 18))
    for a in range(6):
        for b in range(6):
            for ri in range(3):
                for ci in range(3):
                    r = 2 * a + ri if ri < 2 else 12 + a
                    c = 2 * b + ci if ci < 2 else 12 + b
                    out[:, r, c] = A[a, b, ri, ci]
    return out


def _soa_to_fe(Fe, E):
    F = Fe.reshape(6, 3, E)
    out = np.zeros((E, 18))
    for a in range(6):
        out[:, 2 * a] = F[a, 0]
        out[:, 2 * a + 1] = F[a, 1]
        out[:, 12 + a] = F[a, 2]
    return out


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("scheme", ["midpoint", "bdf"])
def test_p2_cell_tensors_match_oracle(lib, scheme):
    mesh, prob = _problem(seed=3)
    u, p, un = T.smooth_fields(prob.x)
    if scheme == "bdf":
        prob.theta, prob.a0 = 1.0, 1.5
        prob.uh = 2.0 * un - 0.5 * (0.8 * un + 0.01)
    _load(lib, prob)
    E, n = prob.cells.shape[0], prob.n
    cells = np.ascontiguousarray(prob.cells, dtype=np.int32)
    x = np.ascontiguousarray(prob.x)
    h = np.ascontiguousarray(prob.h)
    sol = np.concatenate([u, p])
    Ae = np.zeros(324 * E)
    Fe = np.zeros(18 * E)
    lib.p2h_cell_jacobian(E, n, _p(cells), _p(x), _p(h), _p(sol), _p(un), _p(Ae))
    lib.p2h_cell_residual(E, n, _p(cells), _p(x), _p(h), _p(sol), _p(un), _p(Fe))
    A_ref = O.element_matrices(prob, u, p, un)
    assert _rel(_soa_to_ae(Ae, E), A_ref) < TOL
    U, P, Un = O._gather(prob, u, p, un)
    Uh = O._gather_history(prob)
    Fu, _ = PK.element_F(prob, U, P, Un, prob.rules["Fu"], Uh)
    _, Fp = PK.element_F(prob, U, P, Un, prob.rules["Fp"], Uh)
    F_ref = np.concatenate([Fu.reshape(E, 12), Fp], axis=1)
    assert _rel(_soa_to_fe(Fe, E), F_ref) < TOL


def test_p2_shared_rule_is_integrated_once(lib):
    mesh, prob = _problem(seed=1, degrees={k: 8 for k in P2_DEGREE})
    _load(lib, prob)
    assert [lib.p2h_alias(b) for b in range(6)] == [0, 0, 2, 2, 2, 2]
    u, p, un = T.smooth_fields(prob.x)
    E, n = prob.cells.shape[0], prob.n
    Ae = np.zeros(324 * E)
    cells = np.ascontiguousarray(prob.cells, dtype=np.int32)
    sol = np.concatenate([u, p])
    lib.p2h_cell_jacobian(E, n, _p(cells), _p(np.ascontiguousarray(prob.x)), _p(np.ascontiguousarray(prob.h)), _p(sol), _p(un), _p(Ae))
    assert _rel(_soa_to_ae(Ae, E), O.element_matrices(prob, u, p, un)) < TOL


def test_p2_facet_terms_flux_and_schur_operators(lib):
    mesh, prob = _problem(nx=4, ny=4, seed=6)
    _load(lib, prob)
    u, p, un = T.smooth_fields(prob.x)
    E, n = prob.cells.shape[0], prob.n
    ext = M.exterior_facet_indices(mesh.topology)
    fs = O.FacetSet(pairs=mesh.topology.facet_cell_pairs(ext), a_p=1.0, pconst=0.3, a_g=1.0, a_s=0.7, a_n=1.0, beta_n=50.0,
                    a_b=1.0, beta_b=0.2)
    fc, fm = D.pairs_by_cell(fs.pairs)
    co = _Coef(fs.a_p, fs.pconst, fs.a_g, fs.a_s, fs.a_n, fs.beta_n, fs.a_b, fs.beta_b)
    cells = np.ascontiguousarray(prob.cells, dtype=np.int32)
    x = np.ascontiguousarray(prob.x)
    h = np.ascontiguousarray(prob.h)
    sol = np.concatenate([u, p])
    Fe = np.zeros(18 * E)
    Ae = np.zeros(324 * E)
    lib.p2h_facets(0, len(fc), _p(fc), _p(fm), ctypes.byref(co), E, n, _p(cells), _p(x), _p(h), _p(sol), _p(un), _p(Fe))
    lib.p2h_facets(1, len(fc), _p(fc), _p(fm), ctypes.byref(co), E, n, _p(cells), _p(x), _p(h), _p(sol), _p(un), _p(Ae))
    U, P, Un = O._gather(prob, u, p, un)
    ce = fs.pairs[:, 0]
    F_ref = np.zeros((E, 6, 2))
    np.add.at(F_ref, ce, PK.facet_F(prob, fs, U[ce], P[ce], Un[ce]))
    got = _soa_to_fe(Fe, E)[:, :12].reshape(E, 6, 2)
    assert _rel(got, F_ref) < TOL
    A_ref = np.zeros((E, 12, 18))
    np.add.at(A_ref, ce, O.facet_matrices(prob, fs, un))
    assert _rel(_soa_to_ae(Ae, E)[:, :12, :], A_ref) < TOL
    q = lib.p2h_flux(len(fc), _p(fc), _p(fm), _p(cells), _p(x), _p(un))
    assert abs(q - PK.outlet_flux(prob, fs.pairs, un)) < 1e-13
    # stiffness: rows sum to zero, total = int |grad (x + 2y)|^2 for the interpolant of a linear function; mass sums to the area
    Ke = np.zeros(36 * E)
    Me = np.zeros(6 * E)
    lib.p2h_laplace_mass(E, _p(cells), _p(x), _p(Ke), _p(Me))
    K = Ke.reshape(6, 6, E).transpose(2, 0, 1)
    assert np.abs(K.sum(axis=2)).max() < 1e-12 * np.abs(K).max()
    lin = prob.x[:, 0] + 2.0 * prob.x[:, 1]
    energy = np.einsum("ea,eab,eb->", lin[prob.cells], K, lin[prob.cells])
    assert abs(energy - 5.0) < 1e-12                     # |grad|^2 = 5 on the unit square
    assert abs(Me.sum() - 1.0) < 1e-13
