"""CPU arm of bench.py and of the at-size parity tests (TEST / BASELINE INFRASTRUCTURE ONLY): turns a scenario
of the host layer (built with `host_only=True`, i.e. without any CUDA context) into an oracle `Problem` with the same
mesh, parameters, quadrature rules, Dirichlet objects and boundary terms, and marches it with the restated reference
solver configuration (oracle/cpu_reference.CReferenceSolver) or the sparse-LU Newton (oracle/ns_oracle.newton_solve).
"""
from __future__ import annotations

import numpy as np

from . import ns_oracle as O


def oracle_bcs(bcs):
    """('u'|'p', block nodes, values) -> oracle (block, unrolled dofs within the block, g) tuples (2-D)."""
    out = []
    for block, nodes, values in bcs:
        nodes = np.asarray(nodes, dtype=np.int64)
        d = (2 * nodes[:, None] + np.arange(2)[None, :]).reshape(-1) if block == "u" else nodes
        out.append((block, d, np.asarray(values, dtype=float)))
    return out


def problem_from_solver(s, facet_tags=None, tags=None, rules=None):
    """The oracle Problem that mirrors a B200 solver instance after setup(): boundary terms of its variant with the
    setup() multiplicity."""
    from cfd_hemodynamic_b200.fem import mesh as M
    from cfd_hemodynamic_b200.fem import quadrature as Q
    mesh = s.mesh
    cell = mesh.topology.cell_name()
    quad = cell == "quadrilateral"
    if rules is None:
        deg = ({"Fu": 22, "Fp": 20, "uu": 22, "up": 20, "pu": 20, "pp": 18} if quad
               else {"Fu": 12, "Fp": 11, "uu": 12, "up": 11, "pu": 11, "pp": 10})
        rules = {k: (Q.quadrilateral_rule(d) if quad else Q.triangle_rule(d)) for k, d in deg.items()}
    p2 = getattr(s, "p_grade", 1) == 2
    cells = np.ascontiguousarray(s.V.dofmap.list)            # P2: three vertex nodes + three edge nodes per cell
    xn = s.V.tabulate_dof_coordinates()[:, :2].copy()
    if p2:
        rules = {k: Q.triangle_rule(d) for k, d in {"Fu": 20, "Fp": 18, "uu": 20, "up": 18, "pu": 18, "pp": 16}.items()}
    fval = np.asarray(s.f.value, dtype=float).reshape(-1)[:2]
    prob = O.Problem(x=xn, cells=cells, h=mesh.h(2, np.arange(cells.shape[0])),
                     dt=float(s.dt.value), rho=float(s.rho.value), mu=float(s.mu.value), f=fval, rules=rules,
                     facet_rule=Q.interval_gauss(Q.FACET_POINTS_QUAD if quad else 4 if p2 else 2))
    bcs = [("u", bc.block_dofs, bc.g.x.array.copy()) for bc in s.bcu_d]
    bcs += [("p", bc.block_dofs, bc.g.x.array.copy()) for bc in s.bcp_d]
    prob.bcs = oracle_bcs(bcs)
    c = float(s._setup_count)
    pairs = mesh.topology.facet_cell_pairs
    if s.variant == "schur":
        prob.facet_sets = [O.FacetSet(pairs=pairs(M.exterior_facet_indices(mesh.topology)), a_p=1.0, a_g=1.0)]
    elif s.variant == "backflow":
        prob.facet_sets = [O.FacetSet(pairs=pairs(facet_tags.find(tags["outlet"])), a_b=c, beta_b=s.beta_backflow)]
    elif s.variant == "velocity_vascular_backflow":
        prob.facet_sets = [O.FacetSet(pairs=pairs(facet_tags.find(tags["outlet"])), pconst=0.0, a_s=c, a_b=c,
                                      beta_b=s.beta_backflow)]
    elif s.variant == "pressurebc":      # stabilized_schur_pressurebc.py:189-201 on the curl-curl form
        prob.formulation = "curlcurl"
        prob.facet_sets = [O.FacetSet(pairs=pairs(facet_tags.find(tags["inlet"])), pconst=c * s._p_inlet_val, a_n=c,
                                      beta_n=s.beta_nitsche),
                           O.FacetSet(pairs=pairs(facet_tags.find(tags["outlet"])), pconst=c * s._p_outlet_val, a_n=c,
                                      beta_n=s.beta_nitsche)]
    else:
        prob.facet_sets = [O.FacetSet(pairs=pairs(facet_tags.find(tags["inlet"])), pconst=c * s.p_inlet, a_n=c,
                                      beta_n=s.beta_nitsche),
                           O.FacetSet(pairs=pairs(facet_tags.find(tags["outlet"])), pconst=0.0, a_s=c, a_b=c,
                                      beta_b=s.beta_backflow)]
    return prob


class CpuMarcher:
    """Scenario.solve's loop (src/scenario.py:243-307) for the CPU arm: solveStep, the resistance-outlet fixed point
    of the pressure variants (`_update_outlet_pressure`, stabilized_schur_pressure_backflow.py:387-396, Q from the old
    u_prev) and the host's u_prev <- u_sol shift."""

    def __init__(self, scenario, solver="reference", nranks=None, sub_pc="ilu", **tol):
        s = scenario.solver
        self.s = s
        self.prob = problem_from_solver(s, getattr(scenario, "facet_tags", None), getattr(scenario, "tags", None))
        n = self.prob.n
        self.n = n
        self.x = np.concatenate([s.u_prev.x.array.copy(), s.p_prev.x.array.copy()])
        self.un = s.u_prev.x.array.copy()
        self.tol = tol
        self.outlet = None
        if s.variant in ("pressure_backflow", "velocity_vascular_backflow"):
            self.outlet = dict(fs=self.prob.facet_sets[-1], frozen=list(s._p_c_frozen), pc=float(s._p_c))
            self._refresh_outlet()
        self.kind = solver
        if solver == "reference":
            from .cpu_reference import CReferenceSolver
            self.ref = CReferenceSolver(self.prob, nranks=nranks, sub_pc=sub_pc)
        else:
            from .c_oracle import FastAssembler
            self.ref = None
            self.asm = FastAssembler(self.prob) if self.prob.cells.shape[1] == 3 else None

    def _refresh_outlet(self):
        o = self.outlet
        o["fs"].pconst = 0.5 * (sum(o["frozen"]) + o["pc"])

    def step(self):
        s, prob, n = self.s, self.prob, self.n
        if self.ref is not None:
            self.x = self.ref.step(self.x, self.un, **self.tol)
        else:
            x0 = O.remove_nullspace(prob, self.x)
            self.x, its, reason = O.newton_solve(prob, x0, self.un, asm=self.asm, **self.tol)
            if reason < 0:
                raise RuntimeError(f"Did not converge, reason: {reason}.")
        if self.outlet is not None:
            o = self.outlet
            q = O.outlet_flux(prob, o["fs"].pairs, self.un)
            o["pc"] = s.alpha_damping * s.R_resistance * abs(q) + (1.0 - s.alpha_damping) * o["pc"]
            self._refresh_outlet()
        self.un = self.x[:2 * n].copy()
        return self.x
