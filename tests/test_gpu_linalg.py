"""GPU parity of the linear-algebra layer vs scipy on the same inputs."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from tests import common as T

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["triangle", "quadrilateral"])
def setup(request):
    from cfd_hemodynamic_b200._lib import Hemo
    from cfd_hemodynamic_b200.fem import mesh as M
    from oracle import ns_oracle as O
    hemo = Hemo(0)
    nx = 40
    mesh = M.create_unit_square(None, nx, nx, cell_type=request.param)
    prob = T.make_problem(mesh, dt=0.01, rho=1.0, mu=0.05, f=(0.0, 0.0))
    x = prob.x
    n = prob.n
    ext = M.exterior_facet_indices(mesh.topology)
    fsets = [(ext, dict(a_p=1.0, a_g=1.0))]
    walls = np.nonzero(np.isclose(x[:, 0], 0) | np.isclose(x[:, 0], 1) | np.isclose(x[:, 1], 0))[0]
    lidf = M.locate_entities_boundary(mesh, 1, lambda X: np.isclose(X[1], 1.0) & (X[0] > 1e-10) & (X[0] < 1 - 1e-10))
    lid = np.unique(mesh.topology.facet_vertices[lidf])
    g1 = np.zeros(2 * n); g1[0::2] = 1.0
    bcs = [("u", walls, np.zeros(2 * n)), ("u", lid, g1)]
    prob.facet_sets = [O.FacetSet(pairs=mesh.topology.facet_cell_pairs(ext), a_p=1.0, a_g=1.0)]
    prob.bcs = T.oracle_bcs(prob, bcs)
    g, (nrowptr, ncol) = T.setup_gpu(hemo, mesh, prob, fsets, bcs)
    u, p, un = T.smooth_fields(x, seed=2, U=0.5)
    dev = hemo.device
    xd = torch.tensor(np.concatenate([u, p]), device=dev)
    und = torch.tensor(un, device=dev)
    vals = torch.zeros(hemo.nnz, dtype=torch.float64, device=dev)
    hemo.assemble_jacobian(xd, und, vals)
    rowptr, col = hemo.get_pattern()
    A = sp.csr_matrix((vals.cpu().numpy(), col.cpu().numpy(), rowptr.cpu().numpy()), shape=(3 * n, 3 * n))
    unodes = np.unique(np.concatenate([walls, lid]))
    yield dict(hemo=hemo, mesh=mesh, prob=prob, A=A, vals=vals, n=n, nrowptr=nrowptr, ncol=ncol, unodes=unodes, g=g)
    hemo.close()


def test_spmv(setup):
    h, A, n = setup["hemo"], setup["A"], setup["n"]
    rng = np.random.default_rng(0)
    x = rng.standard_normal(3 * n)
    xd = torch.tensor(x, device=h.device)
    yd = torch.empty_like(xd)
    h.spmv(setup["vals"], xd, yd)
    ref = A @ x
    assert np.linalg.norm(yd.cpu().numpy() - ref) <= 1e-13 * np.linalg.norm(ref)


def test_vector_kernels(setup):
    h = setup["hemo"]
    rng = np.random.default_rng(1)
    for N in (1, 31, 1000, 300001):
        x = rng.standard_normal(N); y = rng.standard_normal(N)
        xd = torch.tensor(x, device=h.device); yd = torch.tensor(y, device=h.device)
        assert abs(h.dot(xd, yd) - x @ y) <= 1e-12 * max(1.0, np.linalg.norm(x) * np.linalg.norm(y))
        assert abs(h.norm2(xd) - np.linalg.norm(x)) <= 1e-13 * np.linalg.norm(x)
        h.axpy(-0.75, xd, yd)
        assert np.allclose(yd.cpu().numpy(), y - 0.75 * x, rtol=1e-15, atol=1e-15)
        # reductions are bitwise reproducible
        assert h.dot(xd, yd) == h.dot(xd, yd)


def _solver(setup, **kw):
    from cfd_hemodynamic_b200.linear_solver import BlockSchurSolver
    prob = setup["prob"]
    return BlockSchurSolver(setup["hemo"], setup["nrowptr"], setup["ncol"], setup["unodes"], np.zeros(0, int),
                            dt=prob.dt, rho=prob.rho, mu=prob.mu, project_pressure=True, **kw)


def test_galerkin_product_and_vcycle(setup):
    h, A, n = setup["hemo"], setup["A"], setup["n"]
    s = _solver(setup)
    s.setup(setup["vals"])
    torch.cuda.synchronize()
    A00 = A[:2 * n, :2 * n].tocsr()
    # level-1 operator of the velocity hierarchy == R A00 P with P (x) I2
    lv = s.levels[0][0]
    P2 = sp.kron(lv["P"], sp.eye(2)).tocsr()
    Ac_ref = (P2.T @ A00 @ P2).tocsr()
    Cpat = lv["C"]
    vals1 = h.amg_level_values(0, 1, Cpat.nnz * 4).cpu().numpy().reshape(-1, 2, 2)
    Ac_gpu = sp.bsr_matrix((vals1, Cpat.indices, Cpat.indptr), shape=Ac_ref.shape).tocsr()
    err = abs(Ac_gpu - Ac_ref).max() / abs(Ac_ref).max()
    # the hierarchies are stored in single precision (fp64 accumulation): DESIGN.md §5
    assert err < 5e-7, err
    # V-cycle is a contraction on A00 e = r
    rng = np.random.default_rng(3)
    b = rng.standard_normal(2 * n)
    bd = torch.tensor(b, device=h.device)
    xd = torch.zeros_like(bd)
    h.amg_apply(0, bd, xd, 1)
    r1 = np.linalg.norm(b - A00 @ xd.cpu().numpy()) / np.linalg.norm(b)
    h.amg_apply(0, bd, xd, 3)
    r3 = np.linalg.norm(b - A00 @ xd.cpu().numpy()) / np.linalg.norm(b)
    assert r1 < 0.7 and r3 < r1 * 0.5, (r1, r3)
    # pressure Laplacian hierarchy: lap + lumped mass vs scipy assembly
    lap = s.lap.cpu().numpy()
    L = sp.csr_matrix((lap, setup["ncol"], setup["nrowptr"]), shape=(n, n))
    assert abs(L @ np.ones(n)).max() < 1e-12
    assert abs(s.mass.cpu().numpy().sum() - 1.0) < 1e-13
    bp = rng.standard_normal(n); bp -= bp.mean()
    bpd = torch.tensor(bp, device=h.device)
    xpd = torch.zeros_like(bpd)
    h.amg_apply(1, bpd, xpd, 2)
    xp = xpd.cpu().numpy(); xp -= xp.mean()
    rp = np.linalg.norm(bp - L @ xp) / np.linalg.norm(bp)
    assert rp < 0.5, rp


def test_fgmres_matches_direct_solve(setup):
    h, A, n = setup["hemo"], setup["A"], setup["n"]
    s = _solver(setup, rtol=1e-11, restart=80)
    s.setup(setup["vals"])
    rng = np.random.default_rng(4)
    b = rng.standard_normal(3 * n)
    # consistent rhs for the singular (constant-pressure) system: b in range(A)
    b = A @ rng.standard_normal(3 * n)
    bd = torch.tensor(b, device=h.device)
    yd = torch.zeros_like(bd)
    its, res = s.solve(setup["vals"], bd, yd)
    y = yd.cpu().numpy()
    assert res <= 1e-11
    assert np.linalg.norm(A @ y - b) <= 1e-9 * np.linalg.norm(b)
    assert its < 80, its
    # pressure part has zero mean (null space removed)
    assert abs(y[2 * n:].mean()) < 1e-10 * max(1.0, np.abs(y[2 * n:]).max())


def test_multi_gpu_building_blocks(setup):
    """Per-partition pieces of the multi-GPU driver vs numpy: multi-dot, multi-axpy with
    squared norm, scaling, node masking."""
    h, n = setup["hemo"], setup["n"]
    N = 3 * n
    ldv = (N + 31) // 32 * 32
    rng = np.random.default_rng(7)
    k = 5
    Vh = np.zeros((k, ldv))
    Vh[:, :N] = rng.standard_normal((k, N))
    w = rng.standard_normal(N)
    V = torch.tensor(Vh.reshape(-1), device=h.device)
    wd = torch.tensor(w, device=h.device)
    hh = h.vec_mdot(V, ldv, k, wd)
    assert np.allclose(hh, Vh[:, :N] @ w, rtol=1e-13, atol=1e-13)
    nsq = h.vec_maxpy(V, ldv, hh, -1.0, wd, want_normsq=True)
    ref = w - Vh[:, :N].T @ hh
    assert np.allclose(wd.cpu().numpy(), ref, rtol=1e-13, atol=1e-13)
    assert abs(nsq - ref @ ref) <= 1e-12 * (ref @ ref)
    y = torch.empty_like(wd)
    h.vec_scale(0.25, wd, y)
    assert np.allclose(y.cpu().numpy(), 0.25 * ref)
    mask = np.zeros(n, dtype=np.uint8)
    mask[::3] = 1
    h.mask_nodes(torch.tensor(mask, device=h.device), y)
    yy = y.cpu().numpy()
    assert np.all(yy[:2 * n].reshape(-1, 2)[mask == 1] == 0) and np.all(yy[2 * n:][mask == 1] == 0)
    assert np.allclose(yy[2 * n:][mask == 0], 0.25 * ref[2 * n:][mask == 0])


def test_masked_preconditioner_ignores_ghost_nodes(setup):
    """hemo_set_pc_mask: masked nodes get identity rows in the local A00 and a zero correction."""
    h, n = setup["hemo"], setup["n"]
    s = _solver(setup)
    mask = np.zeros(n, dtype=np.uint8)
    mask[-41:] = 1                                   # the last mesh row plays the ghost layer
    md = torch.tensor(mask, device=h.device)
    h.set_pc_mask(md)
    try:
        s._first = True
        s.setup(setup["vals"])
        rng = np.random.default_rng(9)
        r = torch.tensor(rng.standard_normal(3 * n), device=h.device)
        h.mask_nodes(md, r)
        z = torch.zeros_like(r)
        h.pc_apply(setup["vals"], r, z)
        zu = z.cpu().numpy()[:2 * n].reshape(-1, 2)
        assert np.all(zu[mask == 1] == 0.0)
        assert np.isfinite(z.cpu().numpy()).all() and np.abs(zu[mask == 0]).max() > 0
    finally:
        h.set_pc_mask(None)
