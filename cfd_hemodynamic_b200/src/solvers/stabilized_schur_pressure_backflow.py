"""`--solver stabilized_schur_pressure_backflow` on B200 (reference
src/solvers/stabilized_schur_pressure_backflow.py): weak inlet pressure +
Nitsche tangential penalty (:192-201), resistance outlet p_c = R |Q| with
damped fixed point (:204-209, 387-396) and backflow stabilization (:214-217).

`setup()` appends these terms to F each time it is called (`self.F += ...`);
the reference calls it twice on the `main.py simulate` path, so the terms are
doubled and the first outlet pressure constant stays frozen at R|Q_init|
(SURVEY.md §7.3-1).  The multiplicities are reproduced here.
"""
from typing import Callable

import numpy as np

from ...fem import discretization as D
from ...fem.mesh import QUAD_FACETS
from ._stabilized_common import SET_INLET, SET_OUTLET
from ._stabilized_tet import StabilizedSchurTetB200 as StabilizedSchurB200      # triangles, quadrilaterals and tetrahedra


class Solver(StabilizedSchurB200):
    MAX_ITER = 20
    variant = "pressure_backflow"

    def __init__(self, mesh, dt: float, rho: float, mu: float, f: list,
                 initial_velocity: Callable[[np.ndarray], np.ndarray] = None,
                 p_inlet: float = None, beta_nitsche: float = 100.0, beta_backflow: float = 0.2,
                 R_resistance: float = None, alpha_damping: float = 0.75, p_grade: int = 1, **kwargs):
        if p_inlet is None:
            raise ValueError("p_inlet is required for stabilized_schur_pressure_backflow. "
                             "Pass it via CLI: --p_inlet <value> (in physical units, e.g. Pa)")
        if R_resistance is None:
            raise ValueError("R_resistance is required for stabilized_schur_pressure_backflow. "
                             "Pass it via CLI: --R_resistance <value>")
        self.p_inlet = float(p_inlet)
        self.beta_nitsche = float(beta_nitsche)
        self.beta_backflow = float(beta_backflow)
        self.R_resistance = float(R_resistance)
        self.alpha_damping = float(alpha_damping)
        super().__init__(mesh, dt, rho, mu, f, initial_velocity, p_grade=p_grade, **kwargs)

    # -- boundary terms ------------------------------------------------
    def _inlet_coef(self):
        c = float(self._setup_count)
        return dict(pconst=c * self.p_inlet, a_n=c, beta_n=self.beta_nitsche)

    def _outlet_coef(self):
        c = float(self._setup_count)
        pc_sum = sum(self._p_c_frozen) + self._p_c
        return dict(pconst=0.5 * pc_sum, a_s=c, a_b=c, beta_b=self.beta_backflow)

    def _facet_setup(self, facet_tags, tags):
        fin = facet_tags.find(tags["inlet"])
        fout = facet_tags.find(tags["outlet"])
        self._register_facets(SET_OUTLET, fout, **self._outlet_coef())
        # Q_init from the host u_prev (:204-205); the previous live constant becomes frozen
        if self._setup_count > 1:
            self._p_c_frozen.append(self._p_c)
        if self._host_only:
            q_init = self._host_outlet_flux(fout)
        else:
            q_init = self.hemo.outlet_flux(SET_OUTLET,
                                           self._torch.from_numpy(self.u_prev.x.array).to(self.hemo.device))
        self._p_c = self.R_resistance * abs(q_init)
        self._register_facets(SET_INLET, fin, **self._inlet_coef())
        self._register_facets(SET_OUTLET, fout, **self._outlet_coef())

    def _host_outlet_flux(self, facets) -> float:
        """assemble_scalar(dot(u_prev, n) * ds_out) on the host (host_only mode)."""
        topo = self.mesh.topology
        pairs = topo.facet_cell_pairs(facets)
        x = self.mesh.geometry.x[:, :2]
        cells = self.mesh.geometry.dofmap[pairs[:, 0]]
        lf = pairs[:, 1]
        X = x[cells]
        ar = np.arange(cells.shape[0])
        if cells.shape[1] == 4:
            fv = np.array(QUAD_FACETS)
            inside = X.mean(axis=1)                 # outward = away from the centroid
        else:
            fv = np.array([[1, 2], [0, 2], [0, 1]])
            inside = X[ar, lf]                      # outward = away from the opposite vertex
        va, vb = fv[lf, 0], fv[lf, 1]
        t = X[ar, vb] - X[ar, va]
        nrm = np.stack([t[:, 1], -t[:, 0]], axis=1)
        nrm *= np.sign(np.einsum("ei,ei->e", nrm, 0.5 * (X[ar, va] + X[ar, vb]) - inside))[:, None]
        U = self.u_prev.x.array.reshape(-1, 2)[cells]
        return float(np.sum(np.einsum("ei,ei->e", 0.5 * (U[ar, va] + U[ar, vb]), nrm)))

    def _update_facet_coefs(self):
        self.hemo.set_facet_coef(SET_OUTLET, **self._outlet_coef())

    def _compute_outlet_flux(self) -> float:
        return self.hemo.outlet_flux(SET_OUTLET, self.d_un)

    def _update_outlet_pressure(self) -> None:
        """p_c <- alpha R|Q| + (1-alpha) p_c with Q from u_prev (:387-396).  The
        host shifts u_prev only after solveStep returns (scenario.py:306), so Q
        lags the new solution by one step — reproduced: d_un still holds the
        old u_prev here."""
        q = self._compute_outlet_flux()
        p_new = self.R_resistance * abs(q)
        self._p_c = self.alpha_damping * p_new + (1.0 - self.alpha_damping) * self._p_c
        if self.verbose:
            print(f"  Resistance BC: Q={q:.6e}, p_new={p_new:.4f}, "
                  f"p_damped={self._p_c:.4f} (alpha={self.alpha_damping:.2f})")

    def _after_step(self, device=False):
        self._update_outlet_pressure()
