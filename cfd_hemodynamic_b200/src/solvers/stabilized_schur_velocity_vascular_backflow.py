"""`--solver stabilized_schur_velocity_vascular_backflow` on B200 (reference
src/solvers/stabilized_schur_velocity_vascular_backflow.py): Dirichlet inlet
velocity from the scenario (parabolic profile of peak `v_max`), resistance
outlet p_c = R |Q| with damped fixed point (:185-191, 377-391), viscous
boundary integral and backflow stabilization on the outlet (:192-205).

The outlet model is the one of `stabilized_schur_pressure_backflow` without
the weak inlet-pressure / Nitsche terms; `setup()` appends the outlet terms
to F on every call like there (`self.F += ...`, :191-205), so the
multiplicity rule of SURVEY.md §7.3-1 applies unchanged.
"""
from typing import Callable

import numpy as np

from ._stabilized_common import SET_OUTLET
from .stabilized_schur_pressure_backflow import Solver as _ResistanceOutletSolver


class Solver(_ResistanceOutletSolver):
    MAX_ITER = 20
    variant = "velocity_vascular_backflow"

    def __init__(self, mesh, dt: float, rho: float, mu: float, f: list,
                 initial_velocity: Callable[[np.ndarray], np.ndarray] = None,
                 v_max: float = None, p_grade: int = 1, beta_backflow: float = 0.2,
                 R_resistance: float = None, alpha_damping: float = 0.75, **kwargs):
        if v_max is None:
            raise ValueError("v_max is required for stabilized_schur_velocity_vascular_backflow. "
                             "Pass it via CLI: --v_max <value>")
        if R_resistance is None:
            raise ValueError("R_resistance is required for stabilized_schur_velocity_vascular_backflow. "
                             "Pass it via CLI: --R_resistance <value>")
        self.v_max = float(v_max)
        # p_inlet / beta_nitsche belong to the pressure-inlet variant; no inlet facet set is registered here
        super().__init__(mesh, dt, rho, mu, f, initial_velocity, p_inlet=0.0, beta_nitsche=0.0,
                         beta_backflow=beta_backflow, R_resistance=R_resistance, alpha_damping=alpha_damping,
                         p_grade=p_grade, **kwargs)

    def _facet_setup(self, facet_tags, tags):
        fout = facet_tags.find(tags["outlet"])
        self._register_facets(SET_OUTLET, fout, **self._outlet_coef())
        # Q_init from the host u_prev (:185-187); the previous live constant becomes frozen
        if self._setup_count > 1:
            self._p_c_frozen.append(self._p_c)
        if self._host_only:
            q_init = self._host_outlet_flux(fout)
        else:
            q_init = self.hemo.outlet_flux(SET_OUTLET,
                                           self._torch.from_numpy(self.u_prev.x.array).to(self.hemo.device))
        self._p_c = self.R_resistance * abs(q_init)
        self._register_facets(SET_OUTLET, fout, **self._outlet_coef())
