"""integration/stabilized_schur_b200.py — the file a maintainer drops into the reference tree — driven by duck-typed
DOLFINx stand-ins (tests/fake_dolfinx.py): permuted dof numbering, foreign facet numbering, Dirichlet objects as
unrolled dof lists.  CPU part: the tables handed to the library are the permuted images of the direct ones.
GPU part: the drop-in reproduces the direct solver."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integration"))

from cfd_hemodynamic_b200.fem import generators          # noqa: E402
from tests.fake_dolfinx import FakeBoundaryCondition, FakeDolfinxHost, FakeTags   # noqa: E402

TAGS = {"inlet": 2, "outlet": 3, "wall": 4, "obstacle": None}


def _stenosis_host(**kw):
    import stabilized_schur_b200 as A
    m, ft = generators.stenosis_structured("moderate", cell_type="triangle", res=0.6, L=20.0, x_position_stenosis=8.0)
    host = FakeDolfinxHost(m, seed=3)
    A.DOLFINX_HOOKS["integration_entities"] = host.integration_entities
    Solver = A.make_solver("stabilized_schur_pressure_backflow")
    s = Solver(host, 0.005, 1.06e-3, 3.5e-3, [0.0, 0.0], None, p_inlet=2.0 * 66.661, R_resistance=50.0,
               _host_objects=host.host_objects(), **kw)
    wall_nodes = np.unique(m.topology.facet_vertices[ft.find(4)])
    bcu = [FakeBoundaryCondition(host, wall_nodes, lambda X: np.zeros((2, X.shape[1])))]
    s.setup(bcu, [], facet_tags=FakeTags(host, ft), tags=TAGS)
    return A, m, ft, host, s, wall_nodes


def test_adapter_tables_are_the_permuted_direct_tables():
    A, m, ft, host, s, wall_nodes = _stenosis_host(host_only=True)
    t = s.inner.export_tables()
    # mesh: coordinates in dof order, cells through the dof numbering
    assert np.array_equal(t["cells"], host.perm[m.geometry.dofmap])
    assert np.allclose(t["x"][host.perm], m.geometry.x[:, :2])
    # Dirichlet object: the wall nodes, in dof numbering
    (block, nodes, vals), = t["bcs"]
    assert block == "u" and np.array_equal(np.sort(nodes), np.sort(host.perm[wall_nodes]))
    # facet sets: the (cell, local facet) pairs DOLFINx reports for the tags (cells are numbered alike)
    for sid, tag in ((1, 2), (2, 3)):
        pairs, coef = t["facet_sets"][sid]
        expect = m.topology.facet_cell_pairs(ft.find(tag))
        assert np.array_equal(pairs[np.lexsort(pairs.T[::-1])], expect[np.lexsort(expect.T[::-1])])
    assert t["facet_sets"][1][1]["pconst"] == pytest.approx(2.0 * 66.661)
    assert s.inner._setup_count == 1


def test_module_level_solver_attribute_is_a_class():
    import stabilized_schur_b200 as A
    assert isinstance(A.Solver, type) and A.Solver.B200_VARIANT == "stabilized_schur"


@pytest.mark.gpu
def test_adapter_reproduces_direct_solver_on_gpu():
    from cfd_hemodynamic_b200.src.scenarios.stenosis_pressure_structured import StenosisPressureStructuredSimulation
    tight = dict(snes_rtol=1e-12, snes_stol=0.0, ksp_rtol=1e-11, ksp_restart=120)
    A, m, ft, host, s, _ = _stenosis_host(**tight)
    sc = StenosisPressureStructuredSimulation("stabilized_schur_pressure_backflow", 0.005, 0.02, grade="moderate",
                                              cell_type="triangle", p_inlet=2.0, R_resistance=50.0, res=0.6, L=20.0,
                                              x_position_stenosis=8.0, **tight)
    d = sc.solver
    for _ in range(3):
        s.solveStep()
        s.u_prev.x.array[:] = s.u_sol.x.array[:]          # the host owns the time-level shift (scenario.py:306-307)
        s.p_prev.x.array[:] = s.p_sol.x.array[:]
        d.solveStep()
        d.u_prev.x.array[:] = d.u_sol.x.array[:]
        d.p_prev.x.array[:] = d.p_sol.x.array[:]
    perm = host.perm
    u_ad = s.u_sol.x.array.reshape(-1, 2)[perm].reshape(-1)       # back to geometry-node order
    p_ad = s.p_sol.x.array[perm]
    assert np.linalg.norm(u_ad - d.u_sol.x.array) <= 1e-9 * np.linalg.norm(d.u_sol.x.array)
    assert np.linalg.norm(p_ad - d.p_sol.x.array) <= 1e-9 * np.linalg.norm(d.p_sol.x.array)
    assert s.its_snes == d.its_snes
