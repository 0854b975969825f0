"""GPU parity of the curl-curl / rotational formulation (reference src/solvers/stabilized_schur_pressurebc.py:85-160,
189-201; hemo_set_formulation(HEMO_FORM_CURLCURL), csrc/assembly_curlcurl.cu) through the C-ABI:

* assembled Jacobian and residual with the weak-pressure / curl-form Nitsche facet terms, Dirichlet rows, lifting and set_bc
  against the global oracles (oracle/ns_oracle.py and oracle/ns3d_oracle.py with formulation = "curlcurl") <= 1e-12, on
  perturbed triangles and tetrahedra;
* the `stabilized_schur_pressurebc` plugin on a pressure-driven channel, three time steps against the oracle's LU Newton
  <= 1e-8, setup() called once and twice (the boundary terms double), triangles and tetrahedra."""
import numpy as np
import pytest
import scipy.sparse as sp

from cfd_hemodynamic_b200.fem import discretization as D
from cfd_hemodynamic_b200.fem import mesh as M
from cfd_hemodynamic_b200.fem import quadrature as Q
from oracle import ns3d_oracle as O3
from oracle import ns_oracle as O
from oracle import simplex_oracle as S
from tests import common as T

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

COEF = dict(pconst=0.7, a_n=2.0, beta_n=30.0)


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("with_bc", [False, True])
def test_curlcurl_assembly_triangles(with_bc):
    from cfd_hemodynamic_b200._lib import Hemo
    mesh = T.perturbed_square(9, 7, seed=5)
    prob = T.make_problem(mesh)
    prob.formulation = "curlcurl"
    n = prob.n
    ext = M.exterior_facet_indices(mesh.topology)
    xm = prob.x[mesh.topology.facet_vertices[ext]].mean(axis=1)
    facets = ext[(xm[:, 0] < 1e-9) | (xm[:, 1] > 1 - 1e-9)]       # two sides: corner cells carry two tagged facets
    prob.facet_sets = [O.FacetSet(pairs=mesh.topology.facet_cell_pairs(facets), **COEF)]
    bcs = []
    if with_bc:
        rng = np.random.default_rng(3)
        bottom = np.nonzero(np.isclose(prob.x[:, 1], 0.0))[0]
        right = np.nonzero(np.isclose(prob.x[:, 0], 1.0))[0]
        bcs = [("u", bottom, rng.standard_normal(2 * n)), ("u", right, rng.standard_normal(2 * n))]
        prob.bcs = T.oracle_bcs(prob, bcs)
    hemo = Hemo(0)
    g, _ = T.setup_gpu(hemo, mesh, prob, [(facets, COEF)], bcs or None)
    hemo.set_formulation("curlcurl")
    dev = hemo.device
    u, p, un = T.smooth_fields(prob.x)
    xd = torch.tensor(np.concatenate([u, p]), device=dev)
    und = torch.tensor(un, device=dev)
    vals = torch.zeros(hemo.nnz, dtype=torch.float64, device=dev)
    b = torch.zeros(3 * n, dtype=torch.float64, device=dev)
    hemo.assemble_jacobian(xd, und, vals)
    hemo.assemble_residual(xd, und, g, b)
    rowptr, col = hemo.get_pattern()
    A = sp.csr_matrix((vals.cpu().numpy(), col.cpu().numpy(), rowptr.cpu().numpy()), shape=(3 * n, 3 * n))
    A_ref = O.assemble_J(prob, u, p, un)
    assert np.linalg.norm((A - A_ref).toarray()) <= 1e-12 * np.linalg.norm(A_ref.data)
    b_ref = O.assemble_F(prob, np.concatenate([u, p]), un)
    assert _rel(b.cpu().numpy(), b_ref) < 1e-12
    # the standard form on the same context again: the switch is a pure selector
    hemo.set_formulation("standard")
    hemo.set_facet_coef(0, pconst=0.0, a_n=0.0, beta_n=0.0)
    prob.formulation = "standard"
    prob.facet_sets = []
    hemo.assemble_jacobian(xd, und, vals)
    A2 = sp.csr_matrix((vals.cpu().numpy(), col.cpu().numpy(), rowptr.cpu().numpy()), shape=(3 * n, 3 * n))
    A2_ref = O.assemble_J(prob, u, p, un)
    assert np.linalg.norm((A2 - A2_ref).toarray()) <= 1e-12 * np.linalg.norm(A2_ref.data)
    hemo.close()


@pytest.mark.parametrize("with_bc", [False, True])
def test_curlcurl_assembly_tetrahedra(with_bc):
    from cfd_hemodynamic_b200._lib import Hemo
    from tests.test_gpu_zz_tet3d import _setup
    from tests.test_tet_host import _perturbed_cube
    x, cells = _perturbed_cube(3, seed=4)
    n = x.shape[0]
    h = S.cell_diameter(x, cells)
    rng = np.random.default_rng(11)
    u, p, un = rng.standard_normal((n, 3)), rng.standard_normal(n), rng.standard_normal((n, 3))
    f = np.array([0.3, -0.2, 0.1])
    par = dict(dt=0.01, rho=1.3, mu=0.02)
    rules = {k: S.tet_gauss_jacobi(d) for k, d in dict(Fu=12, Fp=11, uu=12, up=11, pu=11, pp=10).items()}
    frule = S.triangle_facet_rule(4)
    pairs = S.exterior_facets(cells)
    fx = np.array([np.delete(x[cells[c]], lf, axis=0) for c, lf in pairs])
    tagged = np.isclose(fx[:, :, 0], 1.0).all(axis=1) | np.isclose(fx[:, :, 2], 0.0).all(axis=1)
    fpairs = pairs[tagged]
    fcells, fmask = D.pairs_by_cell(fpairs)
    bcs, bc_lists = [], []
    if with_bc:
        gu = rng.standard_normal(3 * n)
        n0 = np.nonzero(np.isclose(x[:, 0], 0.0))[0]
        n1 = np.nonzero(np.isclose(x[:, 1], 0.0))[0]
        bcs = [("u", n0, gu), ("u", n1, 2.0 * gu)]
        udofs = lambda nodes: (3 * nodes[:, None] + np.arange(3)[None]).reshape(-1)
        bc_lists = [udofs(n0), udofs(n1)]
    flag, mult, cellflag, g = D.dirichlet_arrays(n, cells, bcs, gdim=3)
    prob = O3.Problem3D(x=x, cells=cells, f=f, rules=rules, facet_sets=[O.FacetSet(pairs=fpairs, **COEF)],
                        facet_rule=frule, formulation="curlcurl", **par)
    sol = np.concatenate([u.reshape(-1), p])
    A_ref, b_ref = O3.assemble_system(prob, sol, un.reshape(-1), g, bc_lists=bc_lists)
    hemo = Hemo(0)
    Tn, keep = _setup(hemo, x, cells, h, rules, frule, par, f)
    hemo.set_formulation("curlcurl")
    keep["fc"], keep["fm"] = Tn(fcells, torch.int32), Tn(fmask, torch.int32)
    hemo.set_facet_set(0, keep["fc"], keep["fm"], **COEF)
    if with_bc:
        hemo.set_bc(Tn(flag, torch.uint8), Tn(mult), Tn(cellflag, torch.uint8))
    dev = hemo.device
    vals = torch.zeros(hemo.nnz, dtype=torch.float64, device=dev)
    bvec = torch.zeros(4 * n, dtype=torch.float64, device=dev)
    sol_d, un_d, g_d = Tn(sol), Tn(un.reshape(-1)), Tn(g)
    hemo.assemble_jacobian(sol_d, un_d, vals)
    hemo.assemble_residual(sol_d, un_d, g_d if with_bc else None, bvec)
    rowptr, col = hemo.get_pattern()
    torch.cuda.synchronize()
    A_dev = sp.csr_matrix((vals.cpu().numpy(), col.cpu().numpy(), rowptr.cpu().numpy()), shape=(4 * n, 4 * n))
    assert np.linalg.norm((A_dev - A_ref).tocoo().data) < 1e-12 * np.linalg.norm(A_ref.data)
    assert np.linalg.norm(bvec.cpu().numpy() - b_ref) < 1e-12 * np.linalg.norm(b_ref)
    vals2 = torch.zeros_like(vals)
    hemo.assemble_jacobian(sol_d, un_d, vals2)
    assert torch.equal(vals, vals2)                                  # fixed summation order
    hemo.close()


def _channel_2d():
    mesh = M.create_rectangle((0.0, 0.0), (2.0, 1.0), 16, 10)
    x = mesh.geometry.x[:, :2]
    ext = M.exterior_facet_indices(mesh.topology)
    xm = x[mesh.topology.facet_vertices[ext]].mean(axis=1)
    vals = np.where(xm[:, 0] < 1e-9, 2, np.where(xm[:, 0] > 2 - 1e-9, 3, 4)).astype(np.int32)
    return mesh, M.MeshTags(mesh, 1, ext, vals), {"inlet": 2, "outlet": 3, "wall": 4, "obstacle": None}


@pytest.mark.parametrize("double_setup", [False, True])
def test_pressurebc_plugin_triangles_matches_oracle(double_setup):
    from cfd_hemodynamic_b200.fem.space import Function
    from cfd_hemodynamic_b200.src.boundaryCondition import BoundaryCondition
    from cfd_hemodynamic_b200.src.solvers.stabilized_schur_pressurebc import Solver
    from oracle.workload import problem_from_solver
    mesh, ft, tags = _channel_2d()
    tight = dict(snes_rtol=1e-12, snes_atol=1e-10, snes_stol=0.0, ksp_rtol=1e-11, ksp_atol=1e-14, ksp_restart=150)
    with pytest.raises(ValueError):
        Solver(mesh, 0.05, 1.0, 0.1, [0.0, 0.0])
    s = Solver(mesh, 0.05, 1.0, 0.1, [0.0, 0.0], None, p_inlet=2.0, p_outlet=0.4, beta_nitsche=50.0, **tight)
    bcw = BoundaryCondition(Function(s.V))
    bcw.initTopological(1, ft.find(4))
    s.setup([bcw], [], facet_tags=ft, tags=tags)
    if double_setup:
        s.setup([bcw], [], facet_tags=ft, tags=tags)
    assert s.variant == "pressurebc" and s._setup_count == (2 if double_setup else 1)
    prob = problem_from_solver(s, ft, tags)
    assert prob.formulation == "curlcurl" and prob.facet_sets[0].pconst == s._setup_count * 1.0
    n = s.n
    x = np.zeros(3 * n)
    un = np.zeros(2 * n)
    for _ in range(3):
        s.solveStep()
        s.u_prev.x.array[:] = s.u_sol.x.array[:]
        s.p_prev.x.array[:] = s.p_sol.x.array[:]
        x, its, reason = O.newton_solve(prob, x, un, rtol=1e-12, atol=1e-10, stol=0.0)
        assert reason > 0
        un = x[:2 * n].copy()
    assert np.abs(x[:2 * n]).max() > 1e-2                           # the pressure drop drives a flow
    assert _rel(s.u_sol.x.array, x[:2 * n]) < 1e-8
    assert _rel(s.p_sol.x.array, x[2 * n:]) < 1e-8


def test_pressurebc_plugin_tetrahedra_matches_oracle():
    from cfd_hemodynamic_b200.fem.space import Function
    from cfd_hemodynamic_b200.src.boundaryCondition import BoundaryCondition
    from cfd_hemodynamic_b200.src.solvers.stabilized_schur_pressurebc import Solver
    dt, rho, mu = 0.02, 1.0, 0.1
    mesh = M.create_box((0.0, 0.0, 0.0), (2.0, 1.0, 1.0), 6, 3, 3)
    x = mesh.geometry.x
    cells = mesh.geometry.dofmap
    n = x.shape[0]
    ext = M.exterior_facet_indices(mesh.topology)
    xm = x[mesh.topology.facet_vertices[ext]].mean(axis=1)
    vals = np.where(xm[:, 0] < 1e-9, 2, np.where(xm[:, 0] > 2 - 1e-9, 3, 4)).astype(np.int32)
    ft = M.MeshTags(mesh, 2, ext, vals)
    tags = {"inlet": 2, "outlet": 3, "wall": 4, "obstacle": None}
    tight = dict(snes_rtol=1e-12, snes_atol=1e-10, snes_stol=0.0, ksp_rtol=1e-11, ksp_atol=1e-14, ksp_restart=150, amg_cycles_p=2)
    s = Solver(mesh, dt, rho, mu, [0.0, 0.0, 0.0], None, p_inlet=2.0, p_outlet=0.4, beta_nitsche=50.0, **tight)
    bcw = BoundaryCondition(Function(s.V))
    bcw.initTopological(2, ft.find(4))
    s.setup([bcw], [], facet_tags=ft, tags=tags)
    rules = {k: Q.tetrahedron_rule(d) for k, d in dict(Fu=12, Fp=11, uu=12, up=11, pu=11, pp=10).items()}
    pairs = mesh.topology.facet_cell_pairs
    fs_in = O.FacetSet(pairs=pairs(ft.find(2)), pconst=1.0, a_n=1.0, beta_n=50.0)
    fs_out = O.FacetSet(pairs=pairs(ft.find(3)), pconst=0.2, a_n=1.0, beta_n=50.0)
    prob = O3.Problem3D(x=x, cells=cells, dt=dt, rho=rho, mu=mu, f=np.zeros(3), rules=rules, facet_sets=[fs_in, fs_out],
                        facet_rule=Q.triangle_rule(4), formulation="curlcurl")
    wall_nodes = np.unique(mesh.topology.facet_vertices[ft.find(4)])
    prob.bc_dofs = (3 * wall_nodes[:, None] + np.arange(3)[None, :]).reshape(-1)
    xk, un, g = np.zeros(4 * n), np.zeros(3 * n), np.zeros(4 * n)
    for _ in range(2):
        s.solveStep()
        s.u_prev.x.array[:] = s.u_sol.x.array[:]
        s.p_prev.x.array[:] = s.p_sol.x.array[:]
        xk, _ = O3.newton_step(prob, xk, un, g, rtol=1e-12)
        un = xk[:3 * n].copy()
    assert np.abs(xk[:3 * n]).max() > 1e-3
    assert _rel(s.u_sol.x.array, xk[:3 * n]) < 1e-8
    assert _rel(s.p_sol.x.array, xk[3 * n:]) < 1e-8
