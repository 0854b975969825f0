"""End-to-end parity of the hemodynamic solver variants and the DFG scenario vs the
oracle (same meshes, same inputs, tight tolerances on both sides; 1e-8 relative L2)."""
import numpy as np
import pytest

from tests import common as T

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TIGHT = dict(snes_rtol=1e-12, snes_stol=0.0, ksp_rtol=1e-11, ksp_restart=120)


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _march_oracle(prob, x0, un0, steps, after_step=None, nullspace=False):
    from oracle import ns_oracle as O
    xk, un = x0.copy(), un0.copy()
    n = prob.n
    for k in range(steps):
        xk = O.remove_nullspace(prob, xk)           # unconditional, stabilized_schur.py:319
        xk, its, reason = O.newton_solve(prob, xk, un, rtol=1e-12, stol=0.0)
        assert reason > 0, reason
        if after_step:
            after_step(prob, un)
        un = xk[:2 * n].copy()
    return xk


def test_dfg_two_steps_match_oracle():
    from cfd_hemodynamic_b200.src.scenarios.dfg_1 import DFG1Benchmark
    sc = DFG1Benchmark("stabilized_schur", 0.01, 0.02, lc_min=0.05 / 2, lc_max=0.41 / 6, **TIGHT)
    s = sc.solver
    assert not s._nullspace                      # outlet pressure Dirichlet removes the constant mode
    prob = T.oracle_problem_from_solver(s)
    n = s.n
    for _ in range(2):
        s.solveStep()
        s.u_prev.x.array[:] = s.u_sol.x.array[:]
        s.p_prev.x.array[:] = s.p_sol.x.array[:]
    xk = _march_oracle(prob, np.zeros(3 * n), np.zeros(2 * n), 2)
    assert _rel(s.u_sol.x.array, xk[:2 * n]) < 1e-8
    assert _rel(s.p_sol.x.array, xk[2 * n:]) < 1e-8
    cd, cl = sc.drag_lift()
    assert np.isfinite(cd) and np.isfinite(cl) and cd > 0


def test_backflow_variant_matches_oracle():
    from cfd_hemodynamic_b200.src.scenarios.stenosis_mesh_variable import StenosisMeshVariableSimulation
    sc = StenosisMeshVariableSimulation("stabilized_schur_backflow", 0.01, 0.03, grade="moderate", v_max=5.0,
                                        n_elements_radial=3, L=20.0, x_position_stenosis=8.0, **TIGHT)
    s = sc.solver
    prob = T.oracle_problem_from_solver(s, sc.facet_tags, sc.tags)
    n = s.n
    for _ in range(3):
        s.solveStep()
        s.u_prev.x.array[:] = s.u_sol.x.array[:]
        s.p_prev.x.array[:] = s.p_sol.x.array[:]
    xk = _march_oracle(prob, np.zeros(3 * n), np.zeros(2 * n), 3)
    assert _rel(s.u_sol.x.array, xk[:2 * n]) < 1e-8
    assert _rel(s.p_sol.x.array, xk[2 * n:]) < 1e-8


@pytest.mark.parametrize("double_setup", [False, True])
def test_velocity_vascular_backflow_variant_matches_oracle(double_setup):
    """Dirichlet inlet velocity + resistance outlet + backflow stabilization (reference
    stabilized_schur_velocity_vascular_backflow.py:163-205, 377-391): the outlet pressure follows
    p_c <- alpha R |Q| + (1 - alpha) p_c with Q from the previous u_prev; a second setup() doubles
    the outlet terms and freezes the first constant."""
    from cfd_hemodynamic_b200.src.scenarios.stenosis_mesh_variable import StenosisMeshVariableSimulation
    from oracle import ns_oracle as O
    sc = StenosisMeshVariableSimulation("stabilized_schur_velocity_vascular_backflow", 0.01, 0.04, grade="moderate",
                                        v_max=5.0, R_resistance=3.0, n_elements_radial=3, L=20.0,
                                        x_position_stenosis=8.0, **TIGHT)
    if double_setup:
        sc.setup()
    s = sc.solver
    assert s.variant == "velocity_vascular_backflow" and s._setup_count == (2 if double_setup else 1)
    prob = T.oracle_problem_from_solver(s, sc.facet_tags, sc.tags)
    assert len(prob.facet_sets) == 1 and len(prob.bcs) == 2          # outlet terms only; inlet + wall Dirichlet
    n = s.n
    out_pairs = prob.facet_sets[0].pairs
    state = {"frozen": [0.0] * (s._setup_count - 1), "pc": 0.0}      # zero initial velocity: R|Q_init| = 0

    def refresh():
        prob.facet_sets[0].pconst = 0.5 * (sum(state["frozen"]) + state["pc"])

    def after_step(prob_, un_old):
        q = O.outlet_flux(prob_, out_pairs, un_old)
        state["pc"] = s.alpha_damping * s.R_resistance * abs(q) + (1 - s.alpha_damping) * state["pc"]
        refresh()

    refresh()
    steps = 4
    for _ in range(steps):
        s.solveStep()
        s.u_prev.x.array[:] = s.u_sol.x.array[:]
        s.p_prev.x.array[:] = s.p_sol.x.array[:]
    xk = _march_oracle(prob, np.zeros(3 * n), np.zeros(2 * n), steps, after_step)
    assert abs(s._p_c - state["pc"]) <= 1e-9 * max(1.0, abs(state["pc"]))
    assert state["pc"] > 0.0
    assert _rel(s.u_sol.x.array, xk[:2 * n]) < 1e-8
    assert _rel(s.p_sol.x.array, xk[2 * n:]) < 1e-8


@pytest.mark.parametrize("double_setup,cell_type", [(False, "triangle"), (True, "triangle"),
                                                    (False, "quadrilateral"), (True, "quadrilateral")])
def test_pressure_backflow_variant_matches_oracle(double_setup, cell_type):
    """Weak inlet pressure + Nitsche + resistance outlet + backflow; with the second
    setup() call of Simulation.run (simulation.py:269) every boundary term is doubled and
    the first outlet-pressure constant stays frozen."""
    from cfd_hemodynamic_b200.src.scenarios.stenosis_pressure_structured import StenosisPressureStructuredSimulation
    from oracle import ns_oracle as O
    sc = StenosisPressureStructuredSimulation("stabilized_schur_pressure_backflow", 0.005, 0.02, grade="moderate",
                                              cell_type=cell_type, p_inlet=2.0, R_resistance=50.0, res=0.6, L=20.0,
                                              x_position_stenosis=8.0, **TIGHT)
    if double_setup:
        sc.setup()
    s = sc.solver
    assert s._setup_count == (2 if double_setup else 1)
    prob = T.oracle_problem_from_solver(s, sc.facet_tags, sc.tags)
    n = s.n
    out_pairs = prob.facet_sets[1].pairs
    state = {"frozen": [0.0] * (s._setup_count - 1), "pc": 0.0}      # zero initial velocity: R|Q_init| = 0

    def refresh():
        prob.facet_sets[1].pconst = 0.5 * (sum(state["frozen"]) + state["pc"])

    def after_step(prob_, un_old):
        q = O.outlet_flux(prob_, out_pairs, un_old)                   # Q from the *old* u_prev (lag)
        state["pc"] = s.alpha_damping * s.R_resistance * abs(q) + (1 - s.alpha_damping) * state["pc"]
        refresh()

    refresh()
    steps = 4
    for _ in range(steps):
        s.solveStep()
        s.u_prev.x.array[:] = s.u_sol.x.array[:]
        s.p_prev.x.array[:] = s.p_sol.x.array[:]
    xk = _march_oracle(prob, np.zeros(3 * n), np.zeros(2 * n), steps, after_step)
    assert abs(s._p_c - state["pc"]) <= 1e-9 * max(1.0, abs(state["pc"]))
    assert state["pc"] > 0.0                                         # the resistance loop is exercised
    assert _rel(s.u_sol.x.array, xk[:2 * n]) < 1e-8
    assert _rel(s.p_sol.x.array, xk[2 * n:]) < 1e-8
    assert sc.ffr_device() == sc.ffr()                               # FFR probes read from the device state


@pytest.mark.parametrize("cell_type", ["triangle", "quadrilateral"])
def test_golden_case_on_gpu(cell_type):
    """The committed golden vectors (tests/golden) are reproduced by the CUDA path."""
    import os
    import scipy.sparse as sp
    from cfd_hemodynamic_b200._lib import Hemo
    from tests.golden.make_golden import GOLDEN_FILES, build_case
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", GOLDEN_FILES[cell_type]))
    mesh, prob, fsets, bcs, u, p, un = build_case(cell_type)
    for k in prob.rules:
        prob.rules[k] = (gold[f"rule_{k}_pts"], gold[f"rule_{k}_wts"])
    hemo = Hemo(0)
    g, _ = T.setup_gpu(hemo, mesh, prob, fsets, bcs)
    dev = hemo.device
    xd = torch.tensor(np.concatenate([gold["u"], gold["p"]]), device=dev)
    und = torch.tensor(gold["un"], device=dev)
    vals = torch.zeros(hemo.nnz, dtype=torch.float64, device=dev)
    b = torch.zeros(3 * prob.n, dtype=torch.float64, device=dev)
    hemo.assemble_jacobian(xd, und, vals)
    hemo.assemble_residual(xd, und, g, b)
    rowptr, col = hemo.get_pattern()
    N = 3 * prob.n
    # the golden CSR went through SciPy products (Dirichlet zeroing), which drop the explicit
    # zeros the full FE pattern keeps: compare as matrices
    A_gold = sp.csr_matrix((gold["data"], gold["indices"], gold["indptr"]), shape=(N, N))
    A_gpu = sp.csr_matrix((vals.cpu().numpy(), col.cpu().numpy(), rowptr.cpu().numpy()), shape=(N, N))
    d = A_gpu - A_gold
    assert np.sqrt(d.multiply(d).sum()) < 1e-12 * np.sqrt(A_gold.multiply(A_gold).sum())
    assert _rel(b.cpu().numpy(), gold["b"]) < 1e-12
    hemo.close()
