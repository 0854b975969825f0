"""GPU parity: CUDA assembly through the C-ABI vs the numpy oracle on the same
seeded inputs.  Tolerance: 1e-12 relative Frobenius (north_star), written below."""
import numpy as np
import pytest
import scipy.sparse as sp

from tests import common as T

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

REL_TOL = 1e-12


@pytest.fixture(scope="module")
def hemo():
    from cfd_hemodynamic_b200._lib import Hemo
    h = Hemo(0)
    yield h
    h.close()


CELLS = ["triangle", "quadrilateral"]


def _case(hemo, nx, ny, facet_mode, with_bc, seed=0, cell_type="triangle", **pk):
    from cfd_hemodynamic_b200.fem import mesh as M
    from oracle import ns_oracle as O
    mesh = T.perturbed_square(nx, ny, seed=seed, cell_type=cell_type)
    prob = T.make_problem(mesh, **pk)
    ext = M.exterior_facet_indices(mesh.topology)
    x = prob.x
    fsets = []
    if facet_mode == "all":
        fsets = [(ext, dict(a_p=1.0, a_g=1.0))]
    elif facet_mode == "hemo":
        inlet = M.locate_entities_boundary(mesh, 1, lambda X: np.isclose(X[0], 0.0))
        outlet = M.locate_entities_boundary(mesh, 1, lambda X: np.isclose(X[0], 1.0))
        fsets = [(inlet, dict(pconst=2 * 3.7, a_n=2.0, beta_n=100.0)),
                 (outlet, dict(pconst=0.5 * 1.1 + 0.5 * 0.4, a_s=2.0, a_b=2.0, beta_b=0.2))]
    bcs = []
    if with_bc:
        n = prob.n
        walls = np.nonzero(np.isclose(x[:, 1], 0.0) | np.isclose(x[:, 1], 1.0))[0]
        left = np.nonzero(np.isclose(x[:, 0], 0.0))[0]
        rng = np.random.default_rng(5)
        g0 = rng.standard_normal(2 * n)
        g1 = rng.standard_normal(2 * n)
        gp = rng.standard_normal(n)
        bcs = [("u", walls, g0), ("u", left, g1)]      # corner nodes are in both → diagonal 2
        if facet_mode != "hemo":
            right = np.nonzero(np.isclose(x[:, 0], 1.0))[0]
            bcs.append(("p", right, gp))
    prob.facet_sets = [O.FacetSet(pairs=mesh.topology.facet_cell_pairs(f), **c) for f, c in fsets]
    prob.bcs = T.oracle_bcs(prob, bcs)
    g, graph = T.setup_gpu(hemo, mesh, prob, fsets, bcs)
    for sid in range(len(fsets), 8):
        hemo.set_facet_set(sid, None, None)
    if not bcs:
        hemo.set_bc(None, None, None)
    return mesh, prob, g


@pytest.mark.parametrize("cell_type", CELLS)
@pytest.mark.parametrize("facet_mode,with_bc", [("none", False), ("all", False), ("all", True), ("hemo", True)])
def test_jacobian_parity(hemo, facet_mode, with_bc, cell_type):
    from oracle import ns_oracle as O
    mesh, prob, g = _case(hemo, 9, 7, facet_mode, with_bc, cell_type=cell_type)
    u, p, un = T.smooth_fields(prob.x)
    dev = hemo.device
    xd = torch.tensor(np.concatenate([u, p]), device=dev)
    und = torch.tensor(un, device=dev)
    vals = torch.zeros(hemo.nnz, dtype=torch.float64, device=dev)
    hemo.assemble_jacobian(xd, und, vals)
    rowptr, col = hemo.get_pattern()
    torch.cuda.synchronize()
    A_ref = O.assemble_J(prob, u, p, un) if with_bc else O.assemble_J_raw(prob, u, p, un)
    ip, idx = O.sparsity_pattern(prob)
    # sparsity pattern bit-exact
    assert np.array_equal(rowptr.cpu().numpy(), ip)
    assert np.array_equal(col.cpu().numpy(), idx)
    A_gpu = sp.csr_matrix((vals.cpu().numpy(), idx, ip), shape=A_ref.shape)
    diff = (A_gpu - A_ref)
    rel = np.sqrt(diff.multiply(diff).sum()) / np.sqrt(A_ref.multiply(A_ref).sum())
    assert rel < REL_TOL, rel
    if with_bc:
        marker, _, mult = O.bc_arrays(prob)
        d = A_gpu.diagonal()
        assert np.array_equal(d[marker], mult[marker])
        assert mult.max() == 2.0


@pytest.mark.parametrize("cell_type", CELLS)
@pytest.mark.parametrize("facet_mode,with_bc", [("none", False), ("all", True), ("hemo", True)])
def test_residual_parity(hemo, facet_mode, with_bc, cell_type):
    from oracle import ns_oracle as O
    mesh, prob, g = _case(hemo, 8, 11, facet_mode, with_bc, seed=3, cell_type=cell_type)
    u, p, un = T.smooth_fields(prob.x, seed=4)
    dev = hemo.device
    x = np.concatenate([u, p])
    xd = torch.tensor(x, device=dev)
    und = torch.tensor(un, device=dev)
    b = torch.zeros(3 * prob.n, dtype=torch.float64, device=dev)
    hemo.assemble_residual(xd, und, g, b)
    torch.cuda.synchronize()
    b_ref = O.assemble_F(prob, x, un) if with_bc else O.assemble_F_raw(prob, u, p, un)
    rel = np.linalg.norm(b.cpu().numpy() - b_ref) / np.linalg.norm(b_ref)
    assert rel < REL_TOL, rel


@pytest.mark.parametrize("cell_type", CELLS)
def test_zero_previous_velocity_branch(hemo, cell_type):
    """u_n == 0 exercises the eps0 branch of tau_supg1 (stabilized_schur.py:100-103)."""
    from oracle import ns_oracle as O
    mesh, prob, g = _case(hemo, 6, 6, "all", False, mu=1e-3, dt=0.05, cell_type=cell_type)
    u, p, _ = T.smooth_fields(prob.x)
    un = np.zeros_like(u)
    dev = hemo.device
    vals = torch.zeros(hemo.nnz, dtype=torch.float64, device=dev)
    hemo.assemble_jacobian(torch.tensor(np.concatenate([u, p]), device=dev), torch.tensor(un, device=dev), vals)
    A_ref = O.assemble_J_raw(prob, u, p, un)
    rel = np.linalg.norm(vals.cpu().numpy() - A_ref.data) / np.linalg.norm(A_ref.data)
    assert rel < REL_TOL, rel


@pytest.mark.parametrize("cell_type", CELLS)
def test_deterministic(hemo, cell_type):
    mesh, prob, g = _case(hemo, 12, 12, "all", True, cell_type=cell_type)
    u, p, un = T.smooth_fields(prob.x)
    dev = hemo.device
    xd = torch.tensor(np.concatenate([u, p]), device=dev)
    und = torch.tensor(un, device=dev)
    v1 = torch.zeros(hemo.nnz, dtype=torch.float64, device=dev)
    v2 = torch.zeros_like(v1)
    hemo.assemble_jacobian(xd, und, v1)
    hemo.assemble_jacobian(xd, und, v2)
    assert torch.equal(v1, v2)


@pytest.mark.parametrize("cell_type", CELLS)
def test_outlet_flux(hemo, cell_type):
    from cfd_hemodynamic_b200.fem import mesh as M
    from oracle import ns_oracle as O
    mesh, prob, g = _case(hemo, 7, 9, "hemo", True, cell_type=cell_type)
    _, _, un = T.smooth_fields(prob.x)
    outlet = M.locate_entities_boundary(mesh, 1, lambda X: np.isclose(X[0], 1.0))
    q_ref = O.outlet_flux(prob, mesh.topology.facet_cell_pairs(outlet), un)
    q = hemo.outlet_flux(1, torch.tensor(un, device=hemo.device))
    assert abs(q - q_ref) <= 1e-13 * max(1.0, abs(q_ref))


@pytest.mark.parametrize("cell_type", CELLS)
def test_bdf_time_scheme_parity(hemo, cell_type):
    """hemo_set_time_scheme(theta = 1, a0 = 3/2, u_h = 2 u_n - u_nn / 2): the kernels of the
    stabilized_schur_bdf2 variant (reference stabilized_schur_bdf2.py:76-110) vs the oracle,
    Jacobian and residual with facet terms, lifting and Dirichlet rows."""
    from oracle import ns_oracle as O
    mesh, prob, g = _case(hemo, 8, 7, "hemo", True, seed=6, cell_type=cell_type)
    u, p, un = T.smooth_fields(prob.x, seed=8)
    n = prob.n
    unn = un + 0.05 * np.random.default_rng(1).standard_normal(2 * n)
    prob.theta, prob.a0, prob.uh = 1.0, 1.5, 2.0 * un - 0.5 * unn
    dev = hemo.device
    x = np.concatenate([u, p])
    xd = torch.tensor(x, device=dev)
    und = torch.tensor(un, device=dev)
    uhd = torch.tensor(prob.uh, device=dev)
    vals = torch.zeros(hemo.nnz, dtype=torch.float64, device=dev)
    b = torch.zeros(3 * n, dtype=torch.float64, device=dev)
    try:
        hemo.set_time_scheme(1.0, 1.5, uhd)
        hemo.assemble_jacobian(xd, und, vals)
        hemo.assemble_residual(xd, und, g, b)
        torch.cuda.synchronize()
    finally:
        hemo.set_time_scheme(0.5, 1.0, None)
    A_ref = O.assemble_J(prob, u, p, un)
    ip, idx = O.sparsity_pattern(prob)
    A_gpu = sp.csr_matrix((vals.cpu().numpy(), idx, ip), shape=A_ref.shape)
    d = A_gpu - A_ref
    assert np.sqrt(d.multiply(d).sum()) < REL_TOL * np.sqrt(A_ref.multiply(A_ref).sum())
    b_ref = O.assemble_F(prob, x, un)
    assert np.linalg.norm(b.cpu().numpy() - b_ref) < REL_TOL * np.linalg.norm(b_ref)
    # back on the default scheme the mid-point operators are reproduced
    prob.theta, prob.a0, prob.uh = 0.5, 1.0, None
    hemo.assemble_jacobian(xd, und, vals)
    A_ref = O.assemble_J(prob, u, p, un)
    A_gpu = sp.csr_matrix((vals.cpu().numpy(), idx, ip), shape=A_ref.shape)
    d = A_gpu - A_ref
    assert np.sqrt(d.multiply(d).sum()) < REL_TOL * np.sqrt(A_ref.multiply(A_ref).sum())


def test_q1_shared_rule_single_pass(hemo):
    """One rule for every block form: the quadrilateral kernels integrate all blocks in one pass
    (rule aliases) and must still match the oracle."""
    from oracle import ns_oracle as O
    from oracle import q1_oracle as Q1
    mesh = T.perturbed_square(6, 5, seed=2, cell_type="quadrilateral")
    prob = T.make_problem(mesh, rules={k: Q1.tensor_gauss(5) for k in T.BLOCK_ID})
    T.setup_gpu(hemo, mesh, prob)
    for sid in range(8):
        hemo.set_facet_set(sid, None, None)
    hemo.set_bc(None, None, None)
    u, p, un = T.smooth_fields(prob.x)
    dev = hemo.device
    xd = torch.tensor(np.concatenate([u, p]), device=dev)
    und = torch.tensor(un, device=dev)
    vals = torch.zeros(hemo.nnz, dtype=torch.float64, device=dev)
    b = torch.zeros(3 * prob.n, dtype=torch.float64, device=dev)
    hemo.assemble_jacobian(xd, und, vals)
    hemo.assemble_residual(xd, und, None, b)
    A_ref = O.assemble_J_raw(prob, u, p, un)
    assert np.linalg.norm(vals.cpu().numpy() - A_ref.data) < REL_TOL * np.linalg.norm(A_ref.data)
    b_ref = O.assemble_F_raw(prob, u, p, un)
    assert np.linalg.norm(b.cpu().numpy() - b_ref) < REL_TOL * np.linalg.norm(b_ref)


def test_q1_degree26_rules(hemo):
    """14 x 14 / 13 x 13 / 12 x 12 Gauss points (the estimate with the degree of det J added,
    DESIGN.md §4b): rules are a run-time input, the kernels take them unchanged."""
    from oracle import ns_oracle as O
    from oracle import q1_oracle as Q1
    m = dict(Fu=14, Fp=13, uu=14, up=13, pu=13, pp=12)
    mesh = T.perturbed_square(4, 3, seed=5, cell_type="quadrilateral")
    prob = T.make_problem(mesh, rules={k: Q1.tensor_gauss(v) for k, v in m.items()})
    T.setup_gpu(hemo, mesh, prob)
    for sid in range(8):
        hemo.set_facet_set(sid, None, None)
    hemo.set_bc(None, None, None)
    u, p, un = T.smooth_fields(prob.x)
    dev = hemo.device
    xd = torch.tensor(np.concatenate([u, p]), device=dev)
    und = torch.tensor(un, device=dev)
    vals = torch.zeros(hemo.nnz, dtype=torch.float64, device=dev)
    b = torch.zeros(3 * prob.n, dtype=torch.float64, device=dev)
    hemo.assemble_jacobian(xd, und, vals)
    hemo.assemble_residual(xd, und, None, b)
    A_ref = O.assemble_J_raw(prob, u, p, un)
    assert np.linalg.norm(vals.cpu().numpy() - A_ref.data) < REL_TOL * np.linalg.norm(A_ref.data)
    b_ref = O.assemble_F_raw(prob, u, p, un)
    assert np.linalg.norm(b.cpu().numpy() - b_ref) < REL_TOL * np.linalg.norm(b_ref)


def test_q1_laplace_mass(hemo):
    """Pressure Laplacian / lumped mass of the Schur approximation on quadrilaterals: symmetric,
    constants in the kernel, mass sums to the area."""
    import scipy.sparse as sp
    mesh = T.perturbed_square(7, 6, seed=4, cell_type="quadrilateral")
    prob = T.make_problem(mesh)
    _, (nrowptr, ncol) = T.setup_gpu(hemo, mesh, prob)
    lap, mass = hemo.assemble_laplace_mass()
    L = sp.csr_matrix((lap.cpu().numpy(), ncol, nrowptr), shape=(prob.n, prob.n))
    assert abs(L - L.T).max() < 1e-13
    assert np.abs(L @ np.ones(prob.n)).max() < 1e-12
    assert abs(float(mass.sum()) - 1.0) < 1e-13


def test_edge_cases_and_argument_errors(hemo):
    """Single-cell meshes (every vertex on the boundary), empty facet sets, rule-size limits and
    invalid cell types through the C-ABI."""
    from cfd_hemodynamic_b200._lib import HemoError
    from cfd_hemodynamic_b200.fem import mesh as M
    from oracle import ns_oracle as O
    from oracle import q1_oracle as Q1
    for cell_type, X in (("triangle", np.array([[0.0, 0.0], [1.0, 0.1], [0.3, 0.9]])),
                         ("quadrilateral", np.array([[0.0, 0.1], [1.0, 0.0], [0.15, 0.9], [1.2, 1.1]]))):
        nv = X.shape[0]
        mesh = M.Mesh(X, np.arange(nv, dtype=np.int32)[None, :], cell_type=cell_type if nv == 4 else None)
        prob = T.make_problem(mesh)
        ext = M.exterior_facet_indices(mesh.topology)
        assert len(ext) == nv
        fsets = [(ext, dict(a_p=1.0, a_g=1.0, a_s=0.5, a_n=1.0, beta_n=10.0, a_b=1.0, beta_b=0.2, pconst=1.5))]
        prob.facet_sets = [O.FacetSet(pairs=mesh.topology.facet_cell_pairs(f), **c) for f, c in fsets]
        T.setup_gpu(hemo, mesh, prob, fsets)
        for sid in range(1, 8):
            hemo.set_facet_set(sid, None, None)              # m = 0 removes a set
        hemo.set_bc(None, None, None)
        rng = np.random.default_rng(0)
        u, p, un = rng.standard_normal(2 * nv), rng.standard_normal(nv), rng.standard_normal(2 * nv)
        dev = hemo.device
        vals = torch.zeros(hemo.nnz, dtype=torch.float64, device=dev)
        b = torch.zeros(3 * nv, dtype=torch.float64, device=dev)
        hemo.assemble_jacobian(torch.tensor(np.concatenate([u, p]), device=dev), torch.tensor(un, device=dev), vals)
        hemo.assemble_residual(torch.tensor(np.concatenate([u, p]), device=dev), torch.tensor(un, device=dev), None, b)
        A_ref = O.assemble_J_raw(prob, u, p, un)
        assert hemo.nnz == (3 * nv) ** 2 == A_ref.nnz          # dense single-cell pattern
        assert np.linalg.norm(vals.cpu().numpy() - A_ref.data) < REL_TOL * np.linalg.norm(A_ref.data)
        b_ref = O.assemble_F_raw(prob, u, p, un)
        assert np.linalg.norm(b.cpu().numpy() - b_ref) < REL_TOL * np.linalg.norm(b_ref)
        assert hemo.outlet_flux(5, torch.tensor(un, device=dev)) == 0.0   # empty set: zero flux
    # rule-size limits: 256 points on quadrilaterals (context is in quadrilateral mode here), 80 on triangles
    pts, wts = Q1.tensor_gauss(17)                            # 289 points
    with pytest.raises(HemoError):
        hemo.set_quadrature(0, pts, wts)
    assert hemo.lib.hemo_set_cell_type(hemo._ctx, 7) < 0       # HEMO_EINVAL
    assert hemo.lib.hemo_set_time_scheme(hemo._ctx, 0.0, 1.0, None) < 0
    assert hemo.lib.hemo_set_time_scheme(hemo._ctx, 0.5, 1.0, None) == 0
