"""CPU tests of the scenario layer in host_only mode (no CUDA context): mesh generators,
facet tags, Dirichlet tables and boundary-term tables exported for the device."""
import numpy as np
import pytest

from cfd_hemodynamic_b200.fem import generators as G
from cfd_hemodynamic_b200.fem import mesh as M


def test_stenosis_generator_geometry_and_tags():
    m, ft = G.stenosis_structured("severe", res=0.3)
    x = m.geometry.x[:, :2]
    cells = m.geometry.dofmap
    X = x[cells]
    area = 0.5 * ((X[:, 1, 0] - X[:, 0, 0]) * (X[:, 2, 1] - X[:, 0, 1]) - (X[:, 2, 0] - X[:, 0, 0]) * (X[:, 1, 1] - X[:, 0, 1]))
    assert (area > 0).all()
    o = m.mesh_options
    assert abs(x[:, 0].max() - o["L"]) < 1e-12 and abs(x[:, 1].max() - 2 * o["R_in"]) < 1e-12
    # throat radius of the severe grade: R_min = (1 - 0.75) * r_taper(30)
    r_mid = o["R_in"] + (o["R_out"] - o["R_in"]) * 30.0 / o["L"]
    throat = x[np.abs(x[:, 0] - 30.0) < 1e-9]
    assert abs((throat[:, 1].max() - throat[:, 1].min()) / 2 - 0.25 * r_mid) < 1e-9
    # symmetric about the centre line, even inlet count
    n_in = len(ft.find(G.INLET))
    assert n_in % 2 == 0 and n_in == len(ft.find(G.OUTLET))
    assert set(np.unique(ft.values)) == {G.INLET, G.OUTLET, G.WALL}
    assert len(ft.indices) == len(M.exterior_facet_indices(m.topology))


def test_dfg_generator():
    m, ft = G.dfg_cylinder(lc_min=0.05 / 3, lc_max=0.41 / 8)
    x = m.geometry.x[:, :2]
    X = x[m.geometry.dofmap]
    area = 0.5 * np.abs((X[:, 1, 0] - X[:, 0, 0]) * (X[:, 2, 1] - X[:, 0, 1]) - (X[:, 2, 0] - X[:, 0, 0]) * (X[:, 1, 1] - X[:, 0, 1]))
    assert abs(area.sum() - (2.2 * 0.41 - np.pi * 0.05 ** 2)) < 2e-3
    obst = m.topology.facet_vertices[ft.find(G.OBSTACLE)]
    r = np.hypot(x[obst.ravel(), 0] - 0.2, x[obst.ravel(), 1] - 0.2)
    assert np.allclose(r, 0.05, atol=1e-12)
    assert len(ft.find(G.INLET)) > 0 and len(ft.find(G.OUTLET)) > 0 and len(ft.find(G.WALL)) > 0


def test_lid_scenario_tables_host_only():
    from cfd_hemodynamic_b200.src.scenarios.lid_driven2D import LidDriven2DSimulation
    sc = LidDriven2DSimulation("stabilized_schur", 0.01, 0.1, rho=1, mu=0.01, nx=6, host_only=True)
    s = sc.solver
    assert s.hemo is None and s._setup_count == 1
    t = s.export_tables()
    assert t["variant"] == "schur" and t["cells"].shape == (72, 3)
    (b0, n0, g0), (b1, n1, g1) = t["bcs"]
    assert b0 == b1 == "u" and len(n0) == 19 and len(n1) == 5          # walls (3 sides) and open lid interval
    assert np.all(g1[0::2][n1] == 1.0) and np.all(g0 == 0.0)
    pairs, coef = t["facet_sets"][0]
    assert pairs.shape == (24, 2) and coef == {"a_p": 1.0, "a_g": 1.0}
    sc.setup()                                                          # Simulation.run calls setup() again
    assert s._setup_count == 2 and s.export_tables()["facet_sets"][0][1] == {"a_p": 1.0, "a_g": 1.0}


def test_backflow_scenario_doubles_boundary_term_on_second_setup():
    from cfd_hemodynamic_b200.src.scenarios.stenosis_mesh_variable import StenosisMeshVariableSimulation
    sc = StenosisMeshVariableSimulation("stabilized_schur_backflow", 1e-3, 1e-2, grade="moderate", v_max=5.0,
                                        n_elements_radial=2, L=12.0, x_position_stenosis=5.0, host_only=True)
    s = sc.solver
    t1 = s.export_tables()
    assert t1["variant"] == "backflow" and t1["facet_sets"][2][1]["a_b"] == 1.0
    sc.setup()
    t2 = s.export_tables()
    assert t2["facet_sets"][2][1]["a_b"] == 2.0                         # `self.F -= ...` ran twice
    # inlet profile on inlet nodes, walls (listed last) win on the corner nodes
    blocks = [b for b, _, _ in t2["bcs"]]
    assert blocks == ["u", "u"]
    n = s.n
    from cfd_hemodynamic_b200.fem import discretization as D
    flag, mult, cellflag, g = D.dirichlet_arrays(n, t2["cells"], t2["bcs"])
    x = t2["x"]
    corner = np.nonzero((x[:, 0] == 0.0) & ((x[:, 1] == x[:, 1].min()) | (x[:, 1] == x[:, 1].max())))[0]
    assert np.all(g[2 * corner] == 0.0) and np.all(mult[2 * corner] == 2.0)


def test_missing_required_solver_kwargs_raise():
    from cfd_hemodynamic_b200.src.solvers.stabilized_schur_backflow import Solver as B
    from cfd_hemodynamic_b200.src.solvers.stabilized_schur_pressure_backflow import Solver as P
    m = M.create_unit_square(None, 2, 2)
    with pytest.raises(ValueError, match="v_max is required"):
        B(m, 0.01, 1.0, 1.0, [0, 0], host_only=True)
    with pytest.raises(ValueError, match="p_inlet is required"):
        P(m, 0.01, 1.0, 1.0, [0, 0], host_only=True)
    with pytest.raises(ValueError, match="R_resistance is required"):
        P(m, 0.01, 1.0, 1.0, [0, 0], p_inlet=1.0, host_only=True)


def test_partition_overlap_layers():
    from cfd_hemodynamic_b200.parallel import Partition, slab_partition
    m = M.create_unit_square(None, 12, 4)
    x = m.geometry.x[:, :2]
    cells = m.geometry.dofmap
    owner = slab_partition(x[:, 0], 2)
    p1 = Partition(x, cells, owner, 0, overlap=1)
    p3 = Partition(x, cells, owner, 0, overlap=3)
    assert p3.n_owned == p1.n_owned and p3.n_local > p1.n_local
    # with one layer every ghost row is incomplete, with three only the outermost ghosts are
    assert p1.incomplete_mask[p1.n_owned:].all()
    assert 0 < p3.incomplete_mask.sum() < p3.n_local - p3.n_owned
    assert not p3.incomplete_mask[:p3.n_owned].any()


def test_amg_setup_leaves_uncoupled_vertices_to_the_smoother(hemo_lib_built):
    import scipy.sparse as sp
    from cfd_hemodynamic_b200.fem import amg_setup
    n = 300
    main = 2.0 * np.ones(n)
    off = -1.0 * np.ones(n - 1)
    L = sp.diags([off, main, off], [-1, 0, 1]).tolil()
    for i in (50, 51, 52):           # three vertices without couplings
        L[i, :] = 0
        L[:, i] = 0
        L[i, i] = 1.0
    L = L.tocsr()
    L.sort_indices()
    lv = amg_setup.build_hierarchy(L, np.zeros(n, dtype=bool), max_coarse=20)
    P = lv[0]["P"]
    assert P[[50, 51, 52]].nnz == 0
    assert lv[-1]["P"].shape[1] <= 20


def test_velocity_vascular_backflow_tables_host_only():
    """Plugin discovery by name + outlet-only boundary tables with the setup() multiplicity."""
    from cfd_hemodynamic_b200.src.scenarios.stenosis_mesh_variable import StenosisMeshVariableSimulation
    with pytest.raises(RuntimeError, match="R_resistance is required"):      # Scenario wraps ctor errors (scenario.py:101-110)
        StenosisMeshVariableSimulation("stabilized_schur_velocity_vascular_backflow", 1e-3, 1e-2, grade="moderate",
                                       v_max=5.0, n_elements_radial=3, L=20.0, x_position_stenosis=8.0, host_only=True)
    sc = StenosisMeshVariableSimulation("stabilized_schur_velocity_vascular_backflow", 1e-3, 1e-2, grade="moderate",
                                        v_max=5.0, R_resistance=3.0, alpha_damping=0.5, n_elements_radial=3, L=20.0,
                                        x_position_stenosis=8.0, host_only=True)
    s = sc.solver
    assert s.variant == "velocity_vascular_backflow" and s.alpha_damping == 0.5 and s.R_resistance == 3.0
    t = s.export_tables()
    assert sorted(t["facet_sets"]) == [2]                                # outlet set only, no inlet terms
    assert t["facet_sets"][2][1] == dict(pconst=0.0, a_s=1.0, a_b=1.0, beta_b=0.2)
    assert [b for b, _, _ in t["bcs"]] == ["u", "u"]                     # inlet profile + walls
    sc.setup()
    assert s.export_tables()["facet_sets"][2][1]["a_s"] == 2.0 and s._p_c_frozen == [0.0]


def test_quadrilateral_scenario_tables_host_only():
    """stenosis_pressure_structured defaults to the reference's recombined Q1 mesh."""
    from cfd_hemodynamic_b200.src.scenarios.stenosis_pressure import StenosisPressureSimulation
    from cfd_hemodynamic_b200.src.scenarios.stenosis_pressure_structured import StenosisPressureStructuredSimulation
    kw = dict(grade="moderate", p_inlet=2.0, R_resistance=50.0, res=0.6, L=20.0, x_position_stenosis=8.0, host_only=True)
    sq = StenosisPressureStructuredSimulation("stabilized_schur_pressure_backflow", 0.005, 0.02, **kw)
    st = StenosisPressureSimulation("stabilized_schur_pressure_backflow", 0.005, 0.02, **kw)
    assert sq.mesh.topology.cell_name() == "quadrilateral" and st.mesh.topology.cell_name() == "triangle"
    assert sq.solver.n == st.solver.n and 2 * sq.mesh.num_cells == st.mesh.num_cells
    tq = sq.solver.export_tables()
    assert tq["cells"].shape[1] == 4
    # inlet facets are local facet 1 = (0,2) of the first cell column, outlet facets local facet 2 = (1,3)
    assert set(tq["facet_sets"][1][0][:, 1]) == {1} and set(tq["facet_sets"][2][0][:, 1]) == {2}
    # WSS post-processing on quadrilaterals: Couette field u = (y, 0) gives |tau_w| = mu on straight walls
    s = sq.solver
    s.initStressForm()
    x = s.mesh.geometry.x
    u = np.zeros((x.shape[0], 2))
    u[:, 0] = x[:, 1]
    s.u_sol.x.array[:] = u.reshape(-1)
    s.assemble_wss()
    w = s.shear_stress.x.array.reshape(-1, 2)
    row = int(tq["cells"][0][2])                     # vertices 0..row-1 form the bottom wall (slightly tapered)
    straight = (np.arange(x.shape[0]) < row) & (x[:, 0] > 0.5) & (x[:, 0] < 2.0)
    assert straight.any() and np.allclose(np.linalg.norm(w[straight], axis=1), float(s.mu.value), rtol=2e-2)


def test_adaptive_plugin_ramp_host_only():
    """dt ramp bookkeeping of stabilized_schur_adaptive without a device (the solve itself is a GPU test)."""
    from cfd_hemodynamic_b200.src.scenarios.lid_driven2D import LidDriven2DSimulation
    sc = LidDriven2DSimulation("stabilized_schur_adaptive", 0.01, 0.1, rho=1, mu=0.01, nx=4, host_only=True)
    s = sc.solver
    assert s.target_dt == 0.01 and s.step_count_adapt == 0
    s._set_dt(0.002)
    assert float(s.dt.value) == 0.002


# ---- tetrahedra (taylor_green) -------------------------------------------------------------
def test_unit_cube_mesh_and_tetrahedron_rule():
    from math import factorial as fa
    from cfd_hemodynamic_b200.fem import quadrature as Q
    m = M.create_unit_cube(None, 3, 2, 4)
    assert m.topology.cell_name() == "tetrahedron" and m.topology.dim == 3 and m.geometry.dim == 3
    X = m.geometry.x[m.geometry.dofmap]
    det = np.linalg.det(np.stack([X[:, j + 1] - X[:, 0] for j in range(3)], axis=2))
    assert m.num_cells == 6 * 24 and abs(np.abs(det).sum() / 6.0 - 1.0) < 1e-14 and np.abs(det).min() > 1e-3
    ext = M.exterior_facet_indices(m.topology)
    assert len(ext) == 2 * 2 * (3 * 2 + 2 * 4 + 3 * 4)                  # two triangles per boundary square
    # conforming: every interior facet is shared by exactly two cells
    m.topology.create_connectivity(2, 3)
    assert set(np.unique(m.topology._f2c_count)) == {1, 2}
    pts, wts = Q.tetrahedron_rule(12)
    assert len(wts) == 343 and abs(wts.sum() - 1.0 / 6.0) < 1e-15 and (wts > 0).all()
    for a, b, c in ((12, 0, 0), (0, 12, 0), (0, 0, 12), (4, 4, 4), (5, 3, 1)):
        exact = fa(a) * fa(b) * fa(c) / fa(a + b + c + 3)
        assert abs(np.sum(wts * pts[:, 0] ** a * pts[:, 1] ** b * pts[:, 2] ** c) - exact) < 1e-16


def test_taylor_green_scenario_tables_host_only():
    from cfd_hemodynamic_b200.src.scenarios.taylor_green import TaylorGreenSimulation
    sc = TaylorGreenSimulation("stabilized_schur", 0.005, 0.01, rho=1, mu=1.0, n=3, host_only=True)
    s = sc.solver
    n = 4 ** 3
    assert s.hemo is None and s._tet and s.n == n and s.N == 4 * n
    assert s.V.dofmap.index_map_bs == 3 and s.u_prev.x.array.shape == (3 * n,)
    x = sc.mesh.geometry.x
    # initial velocity = exact solution at t = 0, interleaved per node
    assert np.allclose(s.u_prev.x.array.reshape(n, 3), sc.exact_velocity(0)(x.T).T, atol=0, rtol=0)
    boundary = np.nonzero((np.abs(x - 0.5) > 0.5 - 1e-12).any(axis=1))[0]
    (bu,), (bp,) = s.bcu_d, s.bcp_d
    assert np.array_equal(bu.block_dofs, boundary) and np.array_equal(bp.block_dofs, boundary)
    assert len(boundary) == n - 2 ** 3
    du, _ = bu.dof_indices()
    assert np.array_equal(du, (3 * boundary[:, None] + np.arange(3)[None]).reshape(-1))
    # boundary data follow the exact solution in time
    sc.update_boundary_conditions(0.3)
    bu.update(); bp.update()
    assert np.allclose(bu.g.x.array[du], sc.exact_velocity(0.3)(x.T).T.reshape(-1)[du], atol=0, rtol=0)
    assert np.allclose(bp.g.x.array[boundary], sc.exact_pressure(0.3)(x.T)[boundary], atol=0, rtol=0)
    # the all-facet term of stabilized_schur.py:79 is registered on every exterior triangle
    facets, coef = s._facet_tables[0]
    assert len(facets) == 6 * 2 * 9 and coef == {"a_p": 1.0, "a_g": 1.0}
    with pytest.raises(NotImplementedError):
        s.export_tables()


def test_host_postprocessing_on_tetrahedra():
    """l2_norm_sq and the wall-shear-stress vector on P1 tetrahedra (host post-processing of Scenario.solve)."""
    from cfd_hemodynamic_b200.src.scenario import l2_norm_sq
    from cfd_hemodynamic_b200.src.scenarios.taylor_green import TaylorGreenSimulation
    sc = TaylorGreenSimulation("stabilized_schur", 0.005, 0.01, rho=1, mu=0.7, n=2, host_only=True)
    s = sc.solver
    x = sc.mesh.geometry.x
    # int (x + 2y)^2 + z^2 + 1 over the unit cube = 1/3 + 4/3 + 2*1/4*... computed exactly: P1 fields integrate
    # quadratics exactly with the consistent mass matrix
    s.u_sol.x.array[:] = np.stack([x[:, 0] + 2 * x[:, 1], x[:, 2], np.ones(len(x))], axis=1).reshape(-1)
    exact = (1 / 3 + 4 / 3 + 2 * 2 * 0.25) + 1 / 3 + 1.0
    assert abs(l2_norm_sq(sc.mesh, s.u_sol) - exact) < 1e-13
    s.p_sol.x.array[:] = x[:, 0] * 0 + 2.0
    assert abs(l2_norm_sq(sc.mesh, s.p_sol) - 4.0) < 1e-13
    # shear flow u = (gamma z, 0, 0): traction on the wall z = 0 (n = -e_z) is T = -2 mu eps n = (mu gamma, 0, 0),
    # purely tangential; every interior node of that face collects (1/3) T from each of its 6 facets... checked
    # through the sum over the face: sum_i wss_i = (#facets on the face) * T
    gamma = 1.3
    s.u_sol.x.array[:] = np.stack([gamma * x[:, 2], 0 * x[:, 0], 0 * x[:, 0]], axis=1).reshape(-1)
    s.initStressForm()
    s.assemble_wss()
    w = s.shear_stress.x.array.reshape(-1, 3)
    bottom_only = np.isclose(x[:, 2], 0.0) & (np.abs(x[:, :2] - 0.5) < 0.5 - 1e-12).all(axis=1)
    assert bottom_only.sum() == 1
    # the single interior node of the bottom face touches 6 or fewer bottom triangles, each giving T / 3
    T = np.array([0.7 * gamma, 0.0, 0.0])
    k = np.round(w[bottom_only][0, 0] / (T[0] / 3.0))
    assert 4 <= k <= 8 and np.allclose(w[bottom_only][0], k * T / 3.0, atol=1e-14)
