"""Q1 quadrilateral path, CPU checks (no GPU):

* known-answer tests of the Q1 oracle (oracle/q1_oracle.py): J == dF/dx by complex step,
  physical Hessian vs finite differences, constant-pressure null space, patch test;
* the `__host__ __device__` element routines of csrc/q1_element.cuh — the arithmetic the
  CUDA kernels of assembly_q1.cu wrap — compiled with g++ (tests/host_q1) and compared with
  the oracle through the same SoA element buffers and the same gather layout as the device.

Tolerance for the assembled operators: 1e-12 relative Frobenius (north_star)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from cfd_hemodynamic_b200.fem import discretization as D
from cfd_hemodynamic_b200.fem import mesh as M
from cfd_hemodynamic_b200.fem import quadrature as Q
from oracle import ns_oracle as O
from oracle import q1_oracle as Q1
from tests import common as T

HERE = os.path.dirname(os.path.abspath(__file__))
REL_TOL = 1e-12


class _Params(ctypes.Structure):
    _fields_ = [("dt", ctypes.c_double), ("rho", ctypes.c_double), ("mu", ctypes.c_double),
                ("f", ctypes.c_double * 2), ("eps0", ctypes.c_double)]


class _Coef(ctypes.Structure):
    _fields_ = [(k, ctypes.c_double) for k in ("a_p", "pconst", "a_g", "a_s", "a_n", "beta_n", "a_b", "beta_b")]


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(HERE, "host_q1", "q1_host.cpp")
    out_dir = os.path.join(HERE, "host_q1", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libq1host.so")
    deps = [src] + [os.path.join(HERE, "..", "cfd_hemodynamic_b200", "csrc", f) for f in ("q1_element.cuh", "hemo_rules.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src], check=True)
    L = ctypes.CDLL(so)
    L.q1h_flux.restype = ctypes.c_double
    return L


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _problem(nx=6, ny=5, seed=0, **pk):
    mesh = T.perturbed_square(nx, ny, seed=seed, cell_type="quadrilateral")
    prob = T.make_problem(mesh, **pk)
    return mesh, prob


def _load(lib, prob):
    for k, bid in T.BLOCK_ID.items():
        pts, wts = prob.rules[k]
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        wts = np.ascontiguousarray(wts, dtype=np.float64)
        lib.q1h_set_rule(bid, _p(pts), _p(wts), len(wts))
    s, w = (np.ascontiguousarray(a, dtype=np.float64) for a in prob.facet_rule)
    lib.q1h_set_facet_rule(_p(s), _p(w), len(w))
    par = _Params(prob.dt, prob.rho, prob.mu, (ctypes.c_double * 2)(*prob.f), prob.eps0)
    lib.q1h_set_params(ctypes.byref(par))
    lib.q1h_set_time_scheme.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
    uh = None if prob.uh is None else np.ascontiguousarray(prob.uh, dtype=np.float64)
    prob._uh_keepalive = uh
    lib.q1h_set_time_scheme(prob.theta, prob.a0, None if uh is None else _p(uh))


def _gather_matrix(prob, Ae):
    """numpy restatement of k_gather_matrix: Ae is SoA [(a*4+b)*9 + ri*3+ci][E]."""
    E, n = prob.cells.shape[0], prob.n
    A = Ae.reshape(4, 4, 3, 3, E)
    c = prob.cells.astype(np.int64)
    gd = np.stack([2 * c, 2 * c + 1, 2 * n + c], axis=2)            # (E, a, comp) global dof
    rows = np.broadcast_to(gd.transpose(1, 2, 0)[:, None, :, None, :], (4, 4, 3, 3, E))
    cols = np.broadcast_to(gd.transpose(1, 2, 0)[None, :, None, :, :], (4, 4, 3, 3, E))
    M_ = sp.coo_matrix((A.reshape(-1), (rows.reshape(-1), cols.reshape(-1))), shape=(3 * n, 3 * n)).tocsr()
    M_.sort_indices()
    return M_


def _gather_vector(prob, Fe):
    E, n = prob.cells.shape[0], prob.n
    F = Fe.reshape(4, 3, E)
    c = prob.cells.astype(np.int64)
    gd = np.stack([2 * c, 2 * c + 1, 2 * n + c], axis=2).transpose(1, 2, 0)    # (a, comp, E)
    b = np.zeros(3 * n)
    np.add.at(b, gd.reshape(-1), F.reshape(-1))
    return b


def _facet_sets(mesh, mode):
    ext = M.exterior_facet_indices(mesh.topology)
    if mode == "all":
        return [(ext, dict(a_p=1.0, a_g=1.0))]
    if mode == "hemo":
        inlet = M.locate_entities_boundary(mesh, 1, lambda X: np.isclose(X[0], 0.0))
        outlet = M.locate_entities_boundary(mesh, 1, lambda X: np.isclose(X[0], 1.0))
        return [(inlet, dict(pconst=2 * 3.7, a_n=2.0, beta_n=100.0)),
                (outlet, dict(pconst=0.5 * 1.1 + 0.5 * 0.4, a_s=2.0, a_b=2.0, beta_b=0.2))]
    return []


def _rel(A, B):
    d = A - B
    return np.sqrt(d.multiply(d).sum()) / np.sqrt(B.multiply(B).sum())


# ---------------------------------------------------------------- oracle KATs
def test_q1_mesh_topology():
    mesh = M.create_unit_square(None, 4, 3, cell_type="quadrilateral")
    assert mesh.topology.cell_name() == "quadrilateral"
    ext = M.exterior_facet_indices(mesh.topology)
    assert len(ext) == 2 * (4 + 3)
    pairs = mesh.topology.facet_cell_pairs(ext)
    # every exterior facet lies on the boundary of the unit square
    fv = mesh.topology.facet_vertices[ext]
    x = mesh.geometry.x
    mid = 0.5 * (x[fv[:, 0]] + x[fv[:, 1]])
    assert np.all(np.isclose(mid[:, 0], 0) | np.isclose(mid[:, 0], 1) | np.isclose(mid[:, 1], 0) | np.isclose(mid[:, 1], 1))
    # local facet numbering (0,1),(0,2),(1,3),(2,3)
    for (c, lf), f in zip(pairs, ext):
        cv = mesh.geometry.dofmap[c][list(M.QUAD_FACETS[lf])]
        assert sorted(cv) == sorted(mesh.topology.facet_vertices[f])
    h = mesh.h(2, np.arange(mesh.num_cells))
    assert np.allclose(h, np.hypot(0.25, 1 / 3))


def test_q1_hessian_formula():
    mesh, prob = _problem()
    X = prob.x[prob.cells]
    xi, eta, d = 0.3, 0.6, 1e-6
    _, g, _, theta, kappa = Q1.point_geometry(X, xi, eta)
    for dxi, deta in ((d, 0.0), (0.0, d)):
        ph2, g2, *_ = Q1.point_geometry(X, xi + dxi, eta + deta)
        ph1, g1, *_ = Q1.point_geometry(X, xi - dxi, eta - deta)
        dx = np.einsum("a,eai->ei", ph2 - ph1, X)
        Hdx = theta[:, :, None] * np.einsum("eij,ej->ei", kappa, dx)[:, None, :]
        assert np.abs((g2 - g1) - Hdx).max() < 1e-8 * np.abs(Hdx).max()


def test_q1_jacobian_is_derivative_of_residual():
    mesh, prob = _problem(4, 3)
    prob.rules = {k: Q1.tensor_gauss(5) for k in prob.rules}
    ext = M.exterior_facet_indices(mesh.topology)
    prob.facet_sets = [O.FacetSet(pairs=mesh.topology.facet_cell_pairs(ext), a_p=1.0, a_g=1.0, a_s=0.7, a_n=1.0,
                                  beta_n=100.0, a_b=1.0, beta_b=0.2, pconst=3.0)]
    n = prob.n
    rng = np.random.default_rng(2)
    u, p, un = rng.standard_normal(2 * n), rng.standard_normal(n), rng.standard_normal(2 * n)
    A = O.assemble_J_raw(prob, u, p, un).toarray()
    x = np.concatenate([u, p])
    J = np.zeros_like(A)
    for j in range(3 * n):
        xc = x.astype(complex)
        xc[j] += 1e-30j
        J[:, j] = O.assemble_F_raw(prob, xc[:2 * n], xc[2 * n:], un).imag / 1e-30
    assert np.abs(A - J).max() < 1e-13 * np.abs(A).max()


def test_q1_nullspace_and_patch():
    mesh, prob = _problem()
    n = prob.n
    ext = M.exterior_facet_indices(mesh.topology)
    prob.facet_sets = [O.FacetSet(pairs=mesh.topology.facet_cell_pairs(ext), a_p=1.0, a_g=1.0)]
    u, p, un = T.smooth_fields(prob.x)
    A = O.assemble_J_raw(prob, u, p, un)
    v = np.zeros(3 * n)
    v[2 * n:] = 1.0
    assert np.abs(A @ v).max() < 1e-13 * abs(A).max()
    # linear divergence-free velocity + constant pressure: zero viscous/pressure residual on interior rows
    prob.facet_sets = []
    prob.rho = 1e-30
    a = np.array([[0.3, -0.2], [0.5, -0.3]])
    uu = (prob.x @ a).reshape(-1)
    b = O.assemble_F_raw(prob, uu, np.full(n, 2.0), uu)
    interior = (np.abs(prob.x - 0.5) < 0.5 - 1e-12).all(axis=1)
    mask = np.concatenate([np.repeat(interior, 2), interior])
    assert np.abs(b[mask]).max() < 1e-14


def test_q1_reduces_to_triangle_limit_on_affine_cells():
    """On parallelograms theta = 0 and kappa-terms vanish from R for bilinear-free fields:
    for a globally linear velocity the Q1 and P1 residuals of interior nodes both vanish."""
    mesh = M.create_unit_square(None, 4, 4, cell_type="quadrilateral")
    X = mesh.geometry.x[:, :2][mesh.geometry.dofmap]
    _, _, _, theta, _ = Q1.point_geometry(X, 0.37, 0.81)
    assert np.abs(theta - np.array([1.0, -1.0, -1.0, 1.0])[None]).max() < 1e-14


# ---------------------------------------------------------------- element routines vs oracle
@pytest.mark.parametrize("facet_mode", ["none", "all", "hemo"])
def test_q1_element_jacobian_host(lib, facet_mode):
    mesh, prob = _problem(7, 5)
    _load(lib, prob)
    fsets = _facet_sets(mesh, facet_mode)
    prob.facet_sets = [O.FacetSet(pairs=mesh.topology.facet_cell_pairs(f), **c) for f, c in fsets]
    u, p, un = T.smooth_fields(prob.x)
    E, n = prob.cells.shape[0], prob.n
    cells = np.ascontiguousarray(prob.cells, dtype=np.int32)
    x = np.ascontiguousarray(prob.x)
    sol = np.concatenate([u, p])
    Ae = np.zeros(144 * E)
    lib.q1h_cell_jacobian(E, n, _p(cells), _p(x), _p(prob.h), _p(sol), _p(un), _p(Ae))
    for facets, coef in fsets:
        fc, fm = D.facet_set_by_cell(mesh, facets)
        co = _Coef(**{k: coef.get(k, 0.0) for k, _ in _Coef._fields_})
        lib.q1h_facets(1, len(fc), E, n, _p(fc), _p(fm), ctypes.byref(co), _p(cells), _p(x), _p(prob.h), _p(sol),
                       _p(un), None, None, _p(Ae))
    A = _gather_matrix(prob, Ae)
    A_ref = O.assemble_J_raw(prob, u, p, un)
    assert _rel(A, A_ref) < REL_TOL
    ip, idx = O.sparsity_pattern(prob)
    assert np.array_equal(A.indptr, ip) and np.array_equal(A.indices, idx)


@pytest.mark.parametrize("facet_mode,with_bc", [("none", False), ("all", True), ("hemo", True)])
def test_q1_element_residual_host(lib, facet_mode, with_bc):
    mesh, prob = _problem(6, 7, seed=3)
    _load(lib, prob)
    fsets = _facet_sets(mesh, facet_mode)
    prob.facet_sets = [O.FacetSet(pairs=mesh.topology.facet_cell_pairs(f), **c) for f, c in fsets]
    u, p, un = T.smooth_fields(prob.x, seed=4)
    E, n = prob.cells.shape[0], prob.n
    cells = np.ascontiguousarray(prob.cells, dtype=np.int32)
    x = np.ascontiguousarray(prob.x)
    sol = np.concatenate([u, p])
    cellflag = dvec = None
    if with_bc:
        walls = np.nonzero(np.isclose(x[:, 1], 0.0) | np.isclose(x[:, 1], 1.0))[0]
        left = np.nonzero(np.isclose(x[:, 0], 0.0))[0]
        rng = np.random.default_rng(5)
        bcs = [("u", walls, rng.standard_normal(2 * n)), ("u", left, rng.standard_normal(2 * n))]
        if facet_mode != "hemo":
            bcs.append(("p", np.nonzero(np.isclose(x[:, 0], 1.0))[0], rng.standard_normal(n)))
        prob.bcs = T.oracle_bcs(prob, bcs)
        flag, mult, cellflag, g = D.dirichlet_arrays(n, prob.cells, bcs)
        dvec = np.where(flag != 0, g - sol, 0.0)
    Fe = np.zeros(12 * E)
    lib.q1h_cell_residual(E, n, _p(cells), _p(x), _p(prob.h), _p(sol), _p(un),
                          _p(cellflag) if with_bc else None, _p(dvec) if with_bc else None, _p(Fe))
    for facets, coef in fsets:
        fc, fm = D.facet_set_by_cell(mesh, facets)
        co = _Coef(**{k: coef.get(k, 0.0) for k, _ in _Coef._fields_})
        lib.q1h_facets(0, len(fc), E, n, _p(fc), _p(fm), ctypes.byref(co), _p(cells), _p(x), _p(prob.h), _p(sol),
                       _p(un), _p(cellflag) if with_bc else None, _p(dvec) if with_bc else None, _p(Fe))
    b = _gather_vector(prob, Fe)
    if with_bc:
        b[flag != 0] = (sol - g)[flag != 0]
        b_ref = O.assemble_F(prob, sol, un)
    else:
        b_ref = O.assemble_F_raw(prob, u, p, un)
    assert np.linalg.norm(b - b_ref) < REL_TOL * np.linalg.norm(b_ref)


def test_q1_bdf_time_scheme_host(lib):
    """theta = 1, a0 = 3/2, u_h = 2 u_n - u_nn / 2 (stabilized_schur_bdf2.py:76-110): element routines
    vs oracle, Jacobian (cells + facets) and residual with lifting."""
    mesh, prob = _problem(5, 6, seed=8)
    n = prob.n
    rng = np.random.default_rng(12)
    prob.theta, prob.a0 = 1.0, 1.5
    u, p, un = T.smooth_fields(prob.x, seed=6)
    unn = un + 0.05 * rng.standard_normal(2 * n)
    prob.uh = 2.0 * un - 0.5 * unn
    _load(lib, prob)
    fsets = _facet_sets(mesh, "hemo") + _facet_sets(mesh, "all")
    prob.facet_sets = [O.FacetSet(pairs=mesh.topology.facet_cell_pairs(f), **c) for f, c in fsets]
    E = prob.cells.shape[0]
    cells = np.ascontiguousarray(prob.cells, dtype=np.int32)
    x = np.ascontiguousarray(prob.x)
    sol = np.concatenate([u, p])
    walls = np.nonzero(np.isclose(x[:, 1], 0.0) | np.isclose(x[:, 1], 1.0))[0]
    bcs = [("u", walls, rng.standard_normal(2 * n))]
    prob.bcs = T.oracle_bcs(prob, bcs)
    flag, mult, cellflag, g = D.dirichlet_arrays(n, prob.cells, bcs)
    dvec = np.where(flag != 0, g - sol, 0.0)
    Ae, Fe = np.zeros(144 * E), np.zeros(12 * E)
    lib.q1h_cell_jacobian(E, n, _p(cells), _p(x), _p(prob.h), _p(sol), _p(un), _p(Ae))
    lib.q1h_cell_residual(E, n, _p(cells), _p(x), _p(prob.h), _p(sol), _p(un), _p(cellflag), _p(dvec), _p(Fe))
    for facets, coef in fsets:
        fc, fm = D.facet_set_by_cell(mesh, facets)
        co = _Coef(**{k: coef.get(k, 0.0) for k, _ in _Coef._fields_})
        args = (len(fc), E, n, _p(fc), _p(fm), ctypes.byref(co), _p(cells), _p(x), _p(prob.h), _p(sol), _p(un))
        lib.q1h_facets(1, *args, None, None, _p(Ae))
        lib.q1h_facets(0, *args, _p(cellflag), _p(dvec), _p(Fe))
    assert _rel(_gather_matrix(prob, Ae), O.assemble_J_raw(prob, u, p, un)) < REL_TOL
    b = _gather_vector(prob, Fe)
    b[flag != 0] = (sol - g)[flag != 0]
    b_ref = O.assemble_F(prob, sol, un)
    assert np.linalg.norm(b - b_ref) < REL_TOL * np.linalg.norm(b_ref)
    # and the oracle's Jacobian is the derivative of its residual for this scheme
    prob.bcs = []
    prob.rules = {k: Q1.tensor_gauss(4) for k in prob.rules}
    A = O.assemble_J_raw(prob, u, p, un).toarray()
    J = np.zeros_like(A)
    for j in range(3 * n):
        xc = sol.astype(complex)
        xc[j] += 1e-30j
        J[:, j] = O.assemble_F_raw(prob, xc[:2 * n], xc[2 * n:], un).imag / 1e-30
    assert np.abs(A - J).max() < 1e-13 * np.abs(A).max()
    lib.q1h_set_time_scheme(0.5, 1.0, None)


def test_q1_rule_aliases(lib):
    mesh, prob = _problem(3, 3)
    _load(lib, prob)
    # default degrees 22/20/22/20/20/18: F_p own rule; J_up own, J_pu shares J_up's, J_pp own
    assert [lib.q1h_alias(b) for b in range(6)] == [0, 1, 2, 3, 3, 5]
    prob.rules = {k: Q1.tensor_gauss(4) for k in prob.rules}
    _load(lib, prob)
    assert [lib.q1h_alias(b) for b in range(6)] == [0, 0, 2, 2, 2, 2]
    # one shared rule: the single-pass path must agree with the oracle as well
    u, p, un = T.smooth_fields(prob.x)
    E, n = prob.cells.shape[0], prob.n
    cells = np.ascontiguousarray(prob.cells, dtype=np.int32)
    Ae = np.zeros(144 * E)
    lib.q1h_cell_jacobian(E, n, _p(cells), _p(np.ascontiguousarray(prob.x)), _p(prob.h), _p(np.concatenate([u, p])),
                          _p(un), _p(Ae))
    assert _rel(_gather_matrix(prob, Ae), O.assemble_J_raw(prob, u, p, un)) < REL_TOL


def test_q1_flux_and_laplace_host(lib):
    mesh, prob = _problem(5, 6, seed=7)
    _, _, un = T.smooth_fields(prob.x)
    outlet = M.locate_entities_boundary(mesh, 1, lambda X: np.isclose(X[0], 1.0))
    fc, fm = D.facet_set_by_cell(mesh, outlet)
    cells = np.ascontiguousarray(prob.cells, dtype=np.int32)
    x = np.ascontiguousarray(prob.x)
    q = lib.q1h_flux(len(fc), _p(fc), _p(fm), _p(cells), _p(x), _p(un))
    q_ref = O.outlet_flux(prob, mesh.topology.facet_cell_pairs(outlet), un)
    assert abs(q - q_ref) <= 1e-13 * max(1.0, abs(q_ref))
    E = prob.cells.shape[0]
    Ke, Me = np.zeros(16 * E), np.zeros(4 * E)
    lib.q1h_laplace_mass(E, _p(cells), _p(x), _p(Ke), _p(Me))
    K = Ke.reshape(4, 4, E)
    assert np.abs(K.sum(axis=1)).max() < 1e-13            # constants are in the kernel
    assert np.abs(K - K.transpose(1, 0, 2)).max() < 1e-14
    # lumped mass sums to the cell areas (shoelace formula)
    X = prob.x[prob.cells][:, [0, 1, 3, 2]]
    area = 0.5 * np.abs(np.sum(X[:, :, 0] * np.roll(X[:, :, 1], -1, axis=1) - np.roll(X[:, :, 0], -1, axis=1) * X[:, :, 1], axis=1))
    assert np.allclose(Me.reshape(4, E).sum(axis=0), area, rtol=1e-13)
