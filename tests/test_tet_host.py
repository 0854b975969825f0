"""Host-compiled checks of the 3-D device code that follows the tetrahedron cell kernel
(csrc/simplex_element.cuh: simplex_facet; csrc/tet_items.cuh: the per-thread bodies of k_tet_facets,
k_tet_lift, k_gather_matrix3d, k_gather_vector3d, k_spmv_node3d, k_tet_facet_flux) against the oracle.
The items are walked over their index ranges in kernel-launch order by tests/host_simplex/simplex_host.cpp
(g++; test infrastructure only, the library never runs them on the CPU).  Tolerance 1e-12."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from cfd_hemodynamic_b200.fem import discretization as D
from cfd_hemodynamic_b200.fem import quadrature as Q
from oracle import ns3d_oracle as O3
from oracle import ns_oracle as O
from oracle import simplex_oracle as S
from tests import common as T
from tests.test_simplex_oracle import FACET_COEFS

HERE = os.path.dirname(os.path.abspath(__file__))
COEF_KEYS = ("a_p", "pconst", "a_g", "a_s", "a_n", "beta_n", "a_b", "beta_b")


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(HERE, "host_simplex", "simplex_host.cpp")
    out_dir = os.path.join(HERE, "host_simplex", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libsimplexhost.so")
    csrc = os.path.join(HERE, "..", "cfd_hemodynamic_b200", "csrc")
    deps = [src, os.path.join(HERE, "..", "include", "hemo.h")] + [
        os.path.join(csrc, f) for f in ("simplex_element.cuh", "tet_items.cuh", "curlcurl_element.cuh", "hemo_rules.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src], check=True)
    L = ctypes.CDLL(so)
    L.sxh_set_params.argtypes = [ctypes.c_double] * 3 + [ctypes.c_void_p] + [ctypes.c_double] * 3
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dt=np.float64):
    return np.ascontiguousarray(a, dtype=dt)


def _coef8(coef):
    return _c([coef.get(k, 0.0) for k in COEF_KEYS])


def _perturbed_cube(nc, seed=0, amp=0.2):
    x, cells = O3.unit_cube_tets(nc)
    rng = np.random.default_rng(seed)
    interior = (np.abs(x - 0.5) < 0.5 - 1e-12).all(axis=1)
    x = x.copy()
    x[interior] += amp / nc * (rng.random((int(interior.sum()), 3)) - 0.5)
    return x, cells


@pytest.mark.parametrize("coef", FACET_COEFS)
@pytest.mark.parametrize("d", [2, 3])
def test_facet_routine_matches_oracle(lib, d, coef):
    if d == 2:
        prob = T.make_problem(T.perturbed_square(4, 3, seed=2), dt=0.02, rho=1.06, mu=0.035)
        x, cells = prob.x, prob.cells
        frule = (np.asarray(prob.facet_rule[0]).reshape(-1, 1), prob.facet_rule[1])
    else:
        x, cells = _perturbed_cube(2)
        frule = S.triangle_facet_rule(4)
    n = x.shape[0]
    h = S.cell_diameter(x, cells)
    pairs = S.exterior_facets(cells)
    m = len(pairs)
    rng = np.random.default_rng(3)
    u, p, un = rng.standard_normal((n, d)), rng.standard_normal(n), rng.standard_normal((n, d))
    fs = O.FacetSet(pairs=pairs, **coef)
    f3 = np.zeros(3)
    for theta in (0.5, 1.0):
        lib.sxh_set_params(0.02, 1.06, 0.035, _p(f3), O.EPS0, theta, 1.0)
        nv = d + 1
        Fu = np.zeros((m, nv, d))
        J = np.zeros((m, nv, nv, d, d + 1))
        flux = np.zeros(m)
        sol = np.concatenate([u.reshape(-1), p])
        lib.sxh_facets(d, m, n, _p(_c(pairs, np.int32)), _p(_c(cells, np.int32)), _p(_c(x)), _p(_c(h)), _p(sol),
                       _p(_c(un.reshape(-1))), _p(_coef8(coef)), _p(_c(frule[0])), _p(_c(frule[1])), len(frule[1]),
                       _p(Fu), _p(J), _p(flux))
        ce = pairs[:, 0]
        U, P, Un = u[cells][ce], p[cells][ce], un[cells][ce]
        ref = S.facet_F(x, cells, h, pairs, fs, U, P, Un, frule, 1.06, 0.035, theta)
        assert np.abs(Fu - ref).max() < 1e-12 * np.abs(ref).max()
        Jr = S.facet_J(x, cells, h, pairs, fs, Un, frule, 1.06, 0.035, theta)          # (m, d nv, (d+1) nv)
        Jh = np.zeros_like(Jr)
        # J[t, a, b, k, ci] -> row a*d + k, column b*d + ci (ci < d) | d*nv + b (ci = d)
        Jh[:, :, :d * nv] = J[..., :d].transpose(0, 1, 3, 2, 4).reshape(m, d * nv, d * nv)
        Jh[:, :, d * nv:] = J[..., d].transpose(0, 1, 3, 2).reshape(m, d * nv, nv)
        assert np.abs(Jh - Jr).max() < 1e-12 * np.abs(Jr).max()
        assert abs(flux.sum() - S.outlet_flux(x, cells, pairs, un.reshape(-1))) < 1e-13
        top = np.array([np.allclose(np.delete(x[cells[c]], lf, axis=0)[:, d - 1], 1.0) for c, lf in pairs])
        assert abs(flux[top].sum() - S.outlet_flux(x, cells, pairs[top], un.reshape(-1))) < 1e-13


def _gather_tables(cells, nrowptr, ncol, n):
    """What hemo_set_node_graph builds on the device: slot of every (cell, a, b), segments sorted by source."""
    E, nv = cells.shape
    rowof = np.repeat(np.arange(n, dtype=np.int32), np.diff(nrowptr))
    G = sp.csr_matrix((np.arange(len(ncol)) + 1, ncol, nrowptr), shape=(n, n))
    ii = np.repeat(cells, nv, axis=1).reshape(-1)
    jj = np.tile(cells, (1, nv)).reshape(-1)
    slot = np.asarray(G[ii, jj]).reshape(-1).astype(np.int64) - 1
    assert (slot >= 0).all()
    src = np.arange(E * nv * nv)
    order = np.lexsort((src, slot))
    mseg_src = src[order].astype(np.int32)
    mseg_ptr = np.concatenate([[0], np.cumsum(np.bincount(slot, minlength=len(ncol)))]).astype(np.int32)
    vsrc = np.arange(E * nv)
    vnode = cells.reshape(-1)
    vorder = np.lexsort((vsrc, vnode))
    vseg_src = vsrc[vorder].astype(np.int32)
    vseg_ptr = np.concatenate([[0], np.cumsum(np.bincount(vnode, minlength=n))]).astype(np.int32)
    return rowof, mseg_ptr, mseg_src, vseg_ptr, vseg_src


def _pattern3d(nrowptr, ncol, n):
    """CSR pattern in the device layout (k_pattern3d)."""
    nnz_node = len(ncol)
    rowptr = np.zeros(4 * n + 1, dtype=np.int64)
    col = np.zeros(16 * nnz_node, dtype=np.int32)
    for i in range(n):
        r0, deg = nrowptr[i], nrowptr[i + 1] - nrowptr[i]
        nb = ncol[r0:r0 + deg]
        rowc = np.concatenate([(3 * nb[:, None] + np.arange(3)[None]).reshape(-1), 3 * n + nb])
        for k in range(4):
            rs = 12 * r0 + 4 * k * deg if k < 3 else 12 * nnz_node + 4 * r0
            rowptr[3 * i + k if k < 3 else 3 * n + i] = rs
            col[rs:rs + 4 * deg] = rowc
    rowptr[4 * n] = 16 * nnz_node
    return rowptr, col


def _emulate(lib, prob, fpairs, coef, bcs, sol, un, xv, theta=0.5, a0=1.0):
    """Emulated launch sequence on `prob` (ns3d_oracle.Problem3D): (A in the device CSR layout, b, J xv, flux,
    Dirichlet tables)."""
    x, cells = prob.x, prob.cells
    E, n = cells.shape[0], x.shape[0]
    fcells, fmask = D.pairs_by_cell(fpairs)
    flag, mult, cellflag, g = D.dirichlet_arrays(n, cells, bcs, gdim=3)
    nrowptr, ncol = D.node_graph(cells, n)
    rowof, mseg_ptr, mseg_src, vseg_ptr, vseg_src = _gather_tables(cells, nrowptr, ncol, n)
    nnz_node = len(ncol)
    for b, k in enumerate(("Fu", "Fp", "uu", "up", "pu", "pp")):
        pts, wts = prob.rules[k]
        assert lib.sxh_set_rule(3, b, _p(_c(pts)), _p(_c(wts)), len(wts)) == 0
    f3 = _c(prob.f)
    lib.sxh_set_params(prob.dt, prob.rho, prob.mu, _p(f3), prob.eps0, theta, a0)
    Ae, Fe = np.zeros(256 * E), np.zeros(16 * E)
    vals, bvec, y, flux = np.zeros(16 * nnz_node), np.zeros(4 * n), np.zeros(4 * n), np.zeros(1)
    unf, solf, xvf, h = _c(un), _c(sol), _c(xv), _c(prob.h)
    frule = prob.facet_rule
    with_bc = bool(flag.any())
    lib.txh_assemble(E, n, ctypes.c_int64(nnz_node), _p(_c(cells, np.int32)), _p(_c(x)), _p(h), _p(solf), _p(unf), _p(unf),
                     _p(nrowptr), _p(ncol), _p(_c(rowof, np.int32)), _p(mseg_ptr), _p(mseg_src), _p(vseg_ptr), _p(vseg_src),
                     len(fcells), _p(fcells), _p(fmask), _p(_coef8(coef)), _p(_c(frule[0])), _p(_c(frule[1])), len(frule[1]),
                     _p(flag) if with_bc else None, _p(mult), _p(cellflag), _p(g), _p(xvf), _p(Ae), _p(Fe), _p(vals), _p(bvec),
                     _p(y), _p(flux))
    rowptr, col = _pattern3d(nrowptr, ncol, n)
    A_dev = sp.csr_matrix((vals, col, rowptr), shape=(4 * n, 4 * n))
    return A_dev, bvec, y, flux[0], (flag, mult, cellflag, g)


@pytest.mark.parametrize("with_bc", [False, True])
@pytest.mark.parametrize("coef", [FACET_COEFS[0], FACET_COEFS[3]])
def test_emulated_tet_assembly_matches_oracle(lib, coef, with_bc):
    x, cells = _perturbed_cube(3, seed=4)
    E, n = cells.shape[0], x.shape[0]
    rng = np.random.default_rng(11)
    u, p, un = rng.standard_normal((n, 3)), rng.standard_normal(n), rng.standard_normal((n, 3))
    f = np.array([0.3, -0.2, 0.1])
    degs = dict(Fu=6, Fp=5, uu=6, up=5, pu=5, pp=4)          # small rules: the layout, not the quadrature, is under test
    rules = {k: S.tet_gauss_jacobi(v) for k, v in degs.items()}
    pairs = S.exterior_facets(cells)
    # facet set: the faces x = 1 and z = 0 (some cells carry two tagged facets)
    fx = np.array([np.delete(x[cells[c]], lf, axis=0) for c, lf in pairs])                # (m, 3, 3)
    tagged = np.isclose(fx[:, :, 0], 1.0).all(axis=1) | np.isclose(fx[:, :, 2], 0.0).all(axis=1)
    fpairs = pairs[tagged]
    fcells, _ = D.pairs_by_cell(fpairs)
    assert (np.bincount(fpairs[:, 0]).max() == 2) and len(fcells) < len(fpairs)
    # Dirichlet: velocity on x = 0 (one condition) and on y = 0 (a second one: edge dofs get diagonal 2), pressure on x = 1
    bcs, bc_lists = [], []
    if with_bc:
        gu = rng.standard_normal(3 * n)
        gp = rng.standard_normal(n)
        n0 = np.nonzero(np.isclose(x[:, 0], 0.0))[0]
        n1 = np.nonzero(np.isclose(x[:, 1], 0.0))[0]
        n2 = np.nonzero(np.isclose(x[:, 0], 1.0))[0]
        bcs = [("u", n0, gu), ("u", n1, 2.0 * gu), ("p", n2, gp)]
        udofs = lambda nodes: (3 * nodes[:, None] + np.arange(3)[None]).reshape(-1)
        bc_lists = [udofs(n0), udofs(n1), 3 * n + n2]
    prob = O3.Problem3D(x=x, cells=cells, dt=0.01, rho=1.3, mu=0.02, f=f, rules=rules,
                        facet_sets=[O.FacetSet(pairs=fpairs, **coef)], facet_rule=S.triangle_facet_rule(4))
    sol = np.concatenate([u.reshape(-1), p])
    xv = rng.standard_normal(4 * n)
    A_dev, bvec, y, flux, (flag, mult, cellflag, g) = _emulate(lib, prob, fpairs, coef, bcs, sol, un.reshape(-1), xv)
    if with_bc:
        assert mult.max() == 2.0 and cellflag.sum() < E
    A_ref, b_ref = O3.assemble_system(prob, sol, un.reshape(-1), g, bc_lists=bc_lists)
    diff = (A_dev - A_ref).tocoo()
    assert np.linalg.norm(diff.data) < 1e-12 * np.linalg.norm(A_ref.data)
    assert np.linalg.norm(bvec - b_ref) < 1e-12 * np.linalg.norm(b_ref)
    assert np.linalg.norm(y - A_ref @ xv) < 1e-12 * np.linalg.norm(A_ref @ xv)
    assert abs(flux - S.outlet_flux(x, cells, fpairs, un.reshape(-1))) < 1e-13
    if with_bc:
        # Dirichlet rows: unit (or multiplicity) diagonal, x - g on the right-hand side
        d = np.nonzero(flag)[0]
        assert np.array_equal(A_dev.diagonal()[d], mult[d]) and np.allclose(bvec[d], sol[d] - g[d], atol=0, rtol=0)


def test_golden_tet_case_host_emulation(lib):
    """The committed 3-D golden vectors (tests/golden/make_golden_tet.py) are reproduced by the oracle and
    by the host-compiled device code."""
    from tests.golden.make_golden_tet import COEF, GOLDEN_TET, bc_dof_lists, bc_values, build_case
    gold = np.load(os.path.join(HERE, "golden", GOLDEN_TET))
    prob, fpairs, bcs, u, p, un = build_case()
    n = prob.n
    for k in prob.rules:                       # the committed rules: golden pins arithmetic, not tables
        prob.rules[k] = (gold[f"rule_{k}_pts"], gold[f"rule_{k}_wts"])
    prob.facet_rule = (gold["facet_pts"], gold["facet_wts"])
    assert np.array_equal(gold["cells"], prob.cells) and np.array_equal(gold["x"], prob.x)
    assert np.array_equal(gold["fpairs"], fpairs)
    sol = np.concatenate([gold["u"], gold["p"]])
    A_gold = sp.csr_matrix((gold["data"], gold["indices"], gold["indptr"]), shape=(4 * n, 4 * n))
    A, b = O3.assemble_system(prob, sol, gold["un"], bc_values(n, bcs), bc_lists=bc_dof_lists(n, bcs))
    assert np.array_equal(A.indptr, gold["indptr"]) and np.array_equal(A.indices, gold["indices"])
    assert np.linalg.norm(A.data - gold["data"]) <= 1e-13 * np.linalg.norm(gold["data"])
    assert np.linalg.norm(b - gold["b"]) <= 1e-13 * np.linalg.norm(gold["b"])
    A_dev, bvec, _, _, (_, _, _, g) = _emulate(lib, prob, fpairs, COEF, bcs, sol, gold["un"], np.zeros(4 * n))
    assert np.array_equal(g, bc_values(n, bcs))
    assert np.linalg.norm((A_dev - A_gold).tocoo().data) < 1e-12 * np.linalg.norm(gold["data"])
    assert np.linalg.norm(bvec - gold["b"]) < 1e-12 * np.linalg.norm(gold["b"])


@pytest.mark.parametrize("sweeps", [1, 4, 5])
def test_emulated_tet_velocity_solve(lib, sweeps):
    """Items of the first 3-D preconditioner: 3x3 node-diagonal inverses of A00, t_u = r_u - A01 z_p and
    damped block-Jacobi sweeps, on a random matrix stored in the device CSR layout."""
    x, cells = _perturbed_cube(2, seed=1)
    n = x.shape[0]
    nrowptr, ncol = D.node_graph(cells, n)
    nnz_node = len(ncol)
    rowptr, col = _pattern3d(nrowptr, ncol, n)
    rng = np.random.default_rng(2)
    vals = rng.standard_normal(16 * nnz_node)
    A = sp.csr_matrix((vals, col, rowptr), shape=(4 * n, 4 * n))
    A = (A + sp.diags(np.concatenate([8.0 * np.ones(3 * n), np.ones(n)]))).tocsr()      # dominant velocity diagonal
    A.sort_indices()
    assert np.array_equal(A.indices, col)
    vals = np.ascontiguousarray(A.data)
    rows = np.repeat(np.arange(n), np.diff(nrowptr))
    diagslot = np.nonzero(rows == ncol)[0].astype(np.int32)
    ru, zp = rng.standard_normal(3 * n), rng.standard_normal(n)
    dinv, tu, tmp, zu = np.zeros(9 * n), np.zeros(3 * n), np.zeros(3 * n), np.zeros(3 * n)
    omega = 2.0 / 3.0
    lib.txh_velocity_solve(n, ctypes.c_int64(nnz_node), _p(nrowptr), _p(ncol), _p(diagslot), _p(vals), _p(ru), _p(zp), sweeps,
                           ctypes.c_double(omega), _p(dinv), _p(tu), _p(tmp), _p(zu))
    A00, A01 = A[:3 * n, :3 * n].toarray(), A[:3 * n, 3 * n:].toarray()
    Dinv = np.zeros((3 * n, 3 * n))
    for i in range(n):
        Dinv[3 * i:3 * i + 3, 3 * i:3 * i + 3] = np.linalg.inv(A00[3 * i:3 * i + 3, 3 * i:3 * i + 3])
    assert np.abs(dinv.reshape(n, 3, 3) - np.array([Dinv[3 * i:3 * i + 3, 3 * i:3 * i + 3] for i in range(n)])).max() < 1e-12
    t_ref = ru - A01 @ zp
    assert np.abs(tu - t_ref).max() < 1e-12 * np.abs(t_ref).max()
    z = np.zeros(3 * n)
    for _ in range(sweeps):
        z = z + omega * Dinv @ (t_ref - A00 @ z)
    assert np.abs(zu - z).max() < 1e-12 * np.abs(z).max()


def test_emulated_tet_selfp(lib):
    """SELFP on tetrahedra (the reference's Schur approximation, stabilized_schur.py:231-235): entries of
    A11 - A10 diag(A00)^-1 A01 on the distance-2 node graph from the device CSR layout, against SciPy."""
    x, cells = _perturbed_cube(2, seed=3)
    n = x.shape[0]
    nrowptr, ncol = D.node_graph(cells, n)
    nnz_node = len(ncol)
    rowptr, col = _pattern3d(nrowptr, ncol, n)
    rng = np.random.default_rng(8)
    A = sp.csr_matrix((rng.standard_normal(16 * nnz_node), col, rowptr), shape=(4 * n, 4 * n))
    A = (A + sp.diags(np.concatenate([5.0 * np.ones(3 * n), np.zeros(n)]))).tocsr()
    A.sort_indices()
    assert np.array_equal(A.indices, col)
    vals = np.ascontiguousarray(A.data)
    G = sp.csr_matrix((np.ones(nnz_node), ncol, nrowptr), shape=(n, n))
    G2 = (G @ G).tocsr()
    G2.sort_indices()
    rowof2 = np.repeat(np.arange(n, dtype=np.int32), np.diff(G2.indptr))
    col2 = G2.indices.astype(np.int32)
    rows = np.repeat(np.arange(n), np.diff(nrowptr))
    diagslot = np.nonzero(rows == ncol)[0].astype(np.int32)
    out = np.zeros(G2.nnz)
    lib.txh_selfp(ctypes.c_int64(G2.nnz), ctypes.c_int64(nnz_node), _p(rowof2), _p(col2), _p(nrowptr), _p(ncol), _p(diagslot),
                  _p(vals), _p(out))
    N = 3 * n
    A00, A01, A10, A11 = A[:N, :N], A[:N, N:], A[N:, :N], A[N:, N:]
    Sp = (A11 - A10 @ sp.diags(1.0 / A00.diagonal()) @ A01).tocsr()
    got = sp.csr_matrix((out, col2, G2.indptr), shape=(n, n))
    d = (got - Sp).tocoo()
    assert np.abs(d.data).max() < 1e-12 * np.abs(Sp.data).max()


def test_emulated_tet_postprocessing(lib):
    """Device post-processing items on tetrahedra (L2 norm, wall-shear-stress vector) against the host versions
    `Scenario.solve` uses (src/scenario.py:l2_norm_sq, src/solverBase.py:assemble_wss)."""
    from cfd_hemodynamic_b200.fem import mesh as M
    from cfd_hemodynamic_b200.src.scenario import l2_norm_sq
    from cfd_hemodynamic_b200.src.scenarios.taylor_green import TaylorGreenSimulation
    sc = TaylorGreenSimulation("stabilized_schur", 0.005, 0.01, rho=1, mu=0.7, n=3, host_only=True)
    s = sc.solver
    mesh = sc.mesh
    x = np.ascontiguousarray(mesh.geometry.x)
    cells = np.ascontiguousarray(mesh.geometry.dofmap, dtype=np.int32)
    n, E = x.shape[0], cells.shape[0]
    rng = np.random.default_rng(6)
    s.u_sol.x.array[:] = rng.standard_normal(3 * n)
    s.p_sol.x.array[:] = rng.standard_normal(n)
    lib.txh_l2.restype = ctypes.c_double
    lu = lib.txh_l2(E, 3, _p(cells), _p(x), _p(_c(s.u_sol.x.array)))
    lp = lib.txh_l2(E, 1, _p(cells), _p(x), _p(_c(s.p_sol.x.array)))
    assert abs(lu - l2_norm_sq(mesh, s.u_sol)) < 1e-13 * lu and abs(lp - l2_norm_sq(mesh, s.p_sol)) < 1e-13 * lp
    s.initStressForm()
    s.assemble_wss()
    fc, fm = D.facet_set_by_cell(mesh, M.exterior_facet_indices(mesh.topology))
    wss = np.zeros(3 * n)
    sol = np.concatenate([s.u_sol.x.array, s.p_sol.x.array])
    lib.txh_wss(len(fc), _p(fc), _p(fm), _p(cells), _p(x), _p(sol), ctypes.c_double(0.7), _p(wss))
    ref = s.shear_stress.x.array
    assert np.abs(ref).max() > 0 and np.abs(wss - ref).max() < 1e-13 * np.abs(ref).max()
