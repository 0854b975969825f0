"""GPU parity of the P1 tetrahedron element-tensor kernel (hemo_tet_element_tensors) against the
dimension-generic simplex oracle, through the C-ABI.  Tolerance 1e-12 relative (north_star)."""
import numpy as np
import pytest

from oracle import ns_oracle as O
from oracle import simplex_oracle as S

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _cube_tets(n, seed=0, amp=0.2):
    """n^3 cubes split into 6 tetrahedra each (Kuhn), interior vertices perturbed."""
    g = np.linspace(0.0, 1.0, n + 1)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    x = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    idx = lambda i, j, k: (i * (n + 1) + j) * (n + 1) + k
    perms = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
    cells = []
    for i in range(n):
        for j in range(n):
            for k in range(n):
                base = np.array([i, j, k])
                for p in perms:
                    v = [base.copy()]
                    for ax in p:
                        w = v[-1].copy()
                        w[ax] += 1
                        v.append(w)
                    cells.append([idx(*q) for q in v])
    rng = np.random.default_rng(seed)
    interior = (np.abs(x - 0.5) < 0.5 - 1e-12).all(axis=1)
    x[interior] += amp / n * (rng.random((int(interior.sum()), 3)) - 0.5)
    return x, np.asarray(cells, dtype=np.int32)


@pytest.mark.parametrize("theta,a0", [(0.5, 1.0), (1.0, 1.5)])
def test_tet_element_tensors_match_oracle(theta, a0):
    from cfd_hemodynamic_b200._lib import Hemo
    x, cells = _cube_tets(3)
    E, n = cells.shape[0], x.shape[0]
    assert E == 162
    h = S.cell_diameter(x, cells)
    rng = np.random.default_rng(5)
    u = np.stack([np.sin(2.1 * x[:, 0] + 0.3) * np.cos(1.7 * x[:, 1]), -np.cos(1.3 * x[:, 0]) * np.sin(2.3 * x[:, 2] + 0.2),
                  0.5 * np.sin(x[:, 1] + x[:, 2])], axis=1) + 0.01 * rng.standard_normal((n, 3))
    un = 0.9 * u + 0.05 * rng.standard_normal((n, 3))
    p = np.sin(1.1 * x[:, 0]) * x[:, 1] + 0.01 * rng.standard_normal(n)
    uh = 2.0 * un - 0.5 * (un + 0.05 * rng.standard_normal((n, 3))) if theta == 1.0 else un.copy()
    f = np.array([0.3, -0.2, 0.1])
    par = dict(dt=0.01, rho=1.3, mu=0.02, f=f, eps0=O.EPS0, theta=theta, a0=a0)
    rules = [S.tet_gauss_jacobi(deg) for deg in (12, 11, 12, 11, 11, 10)]       # F_u, F_p, J_uu, J_up, J_pu, J_pp
    hemo = Hemo(0)
    dev = hemo.device
    for b, (pts, wts) in enumerate(rules):
        hemo.tet_set_quadrature(b, pts, wts)
    hemo.set_params(par["dt"], par["rho"], par["mu"], f[:2], par["eps0"])
    hemo.set_time_scheme(theta, a0, None)
    T = lambda a, dt=torch.float64: torch.tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    sol = np.concatenate([u.reshape(-1), p])
    Ae, Fe = hemo.tet_element_tensors(T(x), T(cells, torch.int32), T(h), T(sol), T(un.reshape(-1)),
                                      T(uh.reshape(-1)) if theta == 1.0 else None, f)
    torch.cuda.synchronize()
    Ae = Ae.cpu().numpy().reshape(4, 4, 4, 4, E).transpose(4, 0, 1, 2, 3)          # [e, a, b, ri, ci]
    Fe = Fe.cpu().numpy().reshape(4, 4, E).transpose(2, 0, 1)                      # [e, a, comp]
    U, P, Un, Uh = u[cells], p[cells], un[cells], uh[cells]
    kw = dict(Uh=Uh, **par)
    Fu, _ = S.element_F(x, cells, h, U, P, Un, rules[0], **kw)
    _, Fp = S.element_F(x, cells, h, U, P, Un, rules[1], **kw)
    Juu, _, _, _ = S.element_J(x, cells, h, U, P, Un, rules[2], **kw)
    _, Jup, _, _ = S.element_J(x, cells, h, U, P, Un, rules[3], **kw)
    _, _, Jpu, _ = S.element_J(x, cells, h, U, P, Un, rules[4], **kw)
    _, _, _, Jpp = S.element_J(x, cells, h, U, P, Un, rules[5], **kw)
    tol = 1e-12
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert rel(Fe[:, :, :3], Fu) < tol and np.linalg.norm(Fe[:, :, 3] - Fp) < tol * np.linalg.norm(Fu)
    assert rel(Ae[:, :, :, :3, :3], Juu.transpose(0, 1, 3, 2, 4)) < tol
    assert rel(Ae[:, :, :, :3, 3], Jup.transpose(0, 1, 3, 2)) < tol
    assert rel(Ae[:, :, :, 3, :3], Jpu) < tol
    assert rel(Ae[:, :, :, 3, 3], Jpp) < tol
    hemo.close()


def test_tet_csr_assembly_matches_oracle():
    """hemo_set_cell_type(TETRAHEDRON) + the generic entry points: sparsity pattern bit-exact with the
    full FE pattern of [u interleaved (3n) | p (n)], assembled Jacobian and residual (cell integrals)
    <= 1e-12 vs the oracle's element tensors scattered with scipy."""
    import scipy.sparse as sp
    from cfd_hemodynamic_b200._lib import Hemo
    from cfd_hemodynamic_b200.fem import discretization as D
    x, cells = _cube_tets(3, seed=2)
    E, n = cells.shape[0], x.shape[0]
    h = S.cell_diameter(x, cells)
    rng = np.random.default_rng(9)
    u, p, un = rng.standard_normal((n, 3)), rng.standard_normal(n), rng.standard_normal((n, 3))
    f = np.array([0.3, -0.2, 0.1])
    par = dict(dt=0.01, rho=1.3, mu=0.02, f=f, eps0=O.EPS0)
    rules = [S.tet_gauss_jacobi(deg) for deg in (12, 11, 12, 11, 11, 10)]
    hemo = Hemo(0)
    dev = hemo.device
    T = lambda a, dt=torch.float64: torch.tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    hemo.set_mesh(T(x), T(cells, torch.int32), T(h))
    assert hemo.dim == 3
    nrowptr, ncol = D.node_graph(cells, n)
    hemo.set_node_graph(T(nrowptr, torch.int32), T(ncol, torch.int32))
    assert hemo.nnz == 16 * len(ncol)
    for b, (pts, wts) in enumerate(rules):
        hemo.set_quadrature(b, pts, wts)
    hemo.set_params(par["dt"], par["rho"], par["mu"], f[:2], par["eps0"])
    hemo.set_body_force3(f)
    sol = np.concatenate([u.reshape(-1), p])
    vals = torch.zeros(hemo.nnz, dtype=torch.float64, device=dev)
    bvec = torch.zeros(4 * n, dtype=torch.float64, device=dev)
    hemo.assemble_jacobian(T(sol), T(un.reshape(-1)), vals)
    hemo.assemble_residual(T(sol), T(un.reshape(-1)), None, bvec)
    rowptr, col = hemo.get_pattern()
    torch.cuda.synchronize()
    # reference: oracle element tensors scattered with scipy
    U, P, Un = u[cells], p[cells], un[cells]
    Fu, _ = S.element_F(x, cells, h, U, P, Un, rules[0], **par)
    _, Fp = S.element_F(x, cells, h, U, P, Un, rules[1], **par)
    Juu, _, _, _ = S.element_J(x, cells, h, U, P, Un, rules[2], **par)
    _, Jup, _, _ = S.element_J(x, cells, h, U, P, Un, rules[3], **par)
    _, _, Jpu, _ = S.element_J(x, cells, h, U, P, Un, rules[4], **par)
    _, _, _, Jpp = S.element_J(x, cells, h, U, P, Un, rules[5], **par)
    Ae = np.zeros((E, 16, 16))
    Ae[:, :12, :12] = Juu.reshape(E, 12, 12)
    Ae[:, :12, 12:] = Jup.reshape(E, 12, 4)
    Ae[:, 12:, :12] = Jpu.reshape(E, 4, 12)
    Ae[:, 12:, 12:] = Jpp
    c64 = cells.astype(np.int64)
    l2g = np.hstack([(3 * c64[:, :, None] + np.arange(3)[None, None, :]).reshape(-1, 12), 3 * n + c64])
    A_ref = sp.coo_matrix((Ae.reshape(-1), (np.repeat(l2g, 16, axis=1).reshape(-1), np.tile(l2g, (1, 16)).reshape(-1))),
                          shape=(4 * n, 4 * n)).tocsr()
    A_ref.sort_indices()
    assert np.array_equal(rowptr.cpu().numpy(), A_ref.indptr)
    assert np.array_equal(col.cpu().numpy(), A_ref.indices)
    assert np.linalg.norm(vals.cpu().numpy() - A_ref.data) < 1e-12 * np.linalg.norm(A_ref.data)
    b_ref = np.zeros(4 * n)
    np.add.at(b_ref, l2g[:, :12].reshape(-1), Fu.reshape(-1))
    np.add.at(b_ref, l2g[:, 12:].reshape(-1), Fp.reshape(-1))
    assert np.linalg.norm(bvec.cpu().numpy() - b_ref) < 1e-12 * np.linalg.norm(b_ref)
    hemo.close()
