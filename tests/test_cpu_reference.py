"""The CPU arm of bench.py (oracle/cpu_reference.CReferenceSolver: the reference's FGMRES + fieldsplit Schur
FULL/SELFP + GMRES/ASM-ILU(0) configuration restated in C + OpenMP) against the sparse-LU oracle."""
import numpy as np

from cfd_hemodynamic_b200.fem import mesh as M
from oracle import cpu_reference as R
from oracle import ns_oracle as O
from tests import common as T


def _lid(nx):
    mesh = M.create_unit_square(None, nx, nx)
    prob = T.make_problem(mesh, dt=0.01, rho=1.0, mu=0.01, f=(0.0, 0.0))
    x, n = prob.x, prob.n
    ext = M.exterior_facet_indices(mesh.topology)
    prob.facet_sets = [O.FacetSet(pairs=mesh.topology.facet_cell_pairs(ext), a_p=1.0, a_g=1.0)]
    walls = np.nonzero(np.isclose(x[:, 0], 0) | np.isclose(x[:, 0], 1) | np.isclose(x[:, 1], 0))[0]
    lidf = M.locate_entities_boundary(mesh, 1, lambda X: np.isclose(X[1], 1.0) & (X[0] > 1e-10) & (X[0] < 1 - 1e-10))
    lid = np.unique(mesh.topology.facet_vertices[lidf])
    g1 = np.zeros(2 * n)
    g1[0::2] = 1.0
    prob.bcs = T.oracle_bcs(prob, [("u", walls, np.zeros(2 * n)), ("u", lid, g1)])
    return prob


def test_block_pattern_matches_oracle_pattern():
    from cfd_hemodynamic_b200.fem import discretization as D
    prob = _lid(6)
    rowptr, colind = R.block_pattern(*D.node_graph(prob.cells, prob.n))
    rp, ci = O.sparsity_pattern(prob)
    assert np.array_equal(rowptr, rp) and np.array_equal(colind, ci)


def test_c_assembly_matches_numpy_oracle():
    prob = _lid(7)
    S = R.CReferenceSolver(prob, nranks=2)
    u, p, un = T.smooth_fields(prob.x)
    x = np.concatenate([u, p])
    A = S.J(u, p, un)
    A_ref = O.assemble_J(prob, u, p, un)
    assert abs(A - A_ref).max() <= 1e-12 * abs(A_ref).max()
    b = S.F(x, un)
    b_ref = O.assemble_F(prob, x, un)
    assert np.linalg.norm(b - b_ref) <= 1e-12 * np.linalg.norm(b_ref)
    S.close()


def test_reference_configuration_converges_to_the_lu_solution():
    """Two time steps of the lid cavity: the restated PETSc configuration (3 ASM blocks) and the sparse-LU Newton
    agree to the solver tolerance; the constant-pressure null space is detected and projected."""
    prob = _lid(12)
    n = prob.n
    S = R.CReferenceSolver(prob, nranks=3, ksp_rtol=1e-11)
    xk, un = np.zeros(3 * n), np.zeros(2 * n)
    xr, unr = np.zeros(3 * n), np.zeros(2 * n)
    for _ in range(2):
        xk = S.step(xk, un, rtol=1e-12, stol=0.0)
        un = xk[:2 * n].copy()
        xr = O.remove_nullspace(prob, xr)
        xr, its, reason = O.newton_solve(prob, xr, unr, rtol=1e-12, stol=0.0)
        assert reason > 0
        unr = xr[:2 * n].copy()
    assert S._nullspace
    st = S.stats()
    assert st["outer_its"] > 0 and st["inner_its"] > st["outer_its"]
    eu = np.linalg.norm(xk[:2 * n] - xr[:2 * n]) / np.linalg.norm(xr[:2 * n])
    pk, pr = xk[2 * n:] - xk[2 * n:].mean(), xr[2 * n:] - xr[2 * n:].mean()
    assert eu < 1e-8 and np.linalg.norm(pk - pr) / np.linalg.norm(pr) < 1e-8
    S.close()


def test_lu_subsolve_configuration_on_the_pressure_driven_stenosis():
    """The hemodynamic variants ask for `lu` sub-solves (stabilized_schur_pressure_backflow.py:284-288): FGMRES(200) +
    Schur FULL + SELFP with exact A00 / Sp solves needs a handful of outer iterations (ILU(0) blocks stall on this problem)
    and marches to the same solution as the plain sparse-LU Newton, resistance outlet included."""
    import contextlib
    import sys
    from cfd_hemodynamic_b200.src.scenarios.stenosis_pressure_structured import StenosisPressureStructuredSimulation
    from oracle.workload import CpuMarcher
    with contextlib.redirect_stdout(sys.stderr):
        sc = StenosisPressureStructuredSimulation("stabilized_schur_pressure_backflow", 1e-3, 1.0, grade="severe", p_inlet=80.0,
                                                  R_resistance=10.0, res=0.6, cell_type="triangle", host_only=True)
    a = CpuMarcher(sc, solver="reference", nranks=2, sub_pc="lu", rtol=1e-10, stol=0.0)
    b = CpuMarcher(sc, solver="lu", rtol=1e-10, stol=0.0)
    for _ in range(3):
        a.step()
        b.step()
    n = a.n
    st = a.ref.stats()
    assert 0 < st["outer_its"] <= 25 * a.ref.newton_its                     # SELFP with exact sub-solves: O(10) iterations per solve
    assert abs(a.outlet["pc"] - b.outlet["pc"]) <= 1e-6 * max(1.0, abs(b.outlet["pc"]))
    assert np.linalg.norm(a.x[:2 * n] - b.x[:2 * n]) <= 1e-6 * np.linalg.norm(b.x[:2 * n])
    assert np.linalg.norm(a.x[2 * n:] - b.x[2 * n:]) <= 1e-6 * np.linalg.norm(b.x[2 * n:])
