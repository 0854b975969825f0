"""`--solver stabilized_schur_bdf2` on B200 (reference src/solvers/stabilized_schur_bdf2.py):
the stabilized_schur formulation with the spatial terms evaluated at u_sol (fully implicit,
:76-110) and the time derivative (a0 u + a1 u_prev + a2 u_prev2)/dt with updateable BDF
coefficients — BDF1 (1, -1, 0) on the first step, BDF2 (3/2, -2, 1/2) afterwards (:300-310);
`solveStep` saves u_prev into u_prev2 after the solve (:323-327).

On the device this is the same set of kernels with another time scheme
(`hemo_set_time_scheme`: theta = 1, a0, history vector u_h = -(a1 u_prev + a2 u_prev2)); the
stabilization parameters still take u_prev (:101-103).
"""
from typing import Callable

import numpy as np

from ...fem.space import Function
from ._stabilized_common import SET_ALL
from ._stabilized_tet import StabilizedSchurTetB200 as StabilizedSchurB200      # triangles, quadrilaterals and tetrahedra


class Solver(StabilizedSchurB200):
    MAX_ITER = 20
    variant = "schur"          # same boundary terms as stabilized_schur (all-facet term, :89)

    def __init__(self, mesh, dt: float, rho: float, mu: float, f: list,
                 initial_velocity: Callable[[np.ndarray], np.ndarray] = None, **kwargs):
        self.step_count = 0
        self.bdf_a0, self.bdf_a1, self.bdf_a2 = 1.0, -1.0, 0.0
        self.d_un2 = self.d_uh = None
        super().__init__(mesh, dt, rho, mu, f, initial_velocity, **kwargs)
        self.u_prev2 = Function(self.V)                       # u at time n-1 (:67)
        if self._pc_kw.get("schur_mode", "laplace") == "laplace":
            # S ~ B (a0 rho/dt M + mu K)^-1 B^T for the fully implicit scheme
            self._pc_kw.setdefault("schur_lap_coef", 1.5 * float(rho) / float(dt))
        if self.hemo is not None:
            torch = self._torch
            self.d_un2 = torch.zeros_like(self.d_un)
            self.d_uh = torch.zeros_like(self.d_un)

    def _prepare_time_scheme(self):
        """bdf_a0/a1/a2 of this step (:301-309) and the history vector on the device."""
        if self.step_count == 0:
            self.bdf_a0, self.bdf_a1, self.bdf_a2 = 1.0, -1.0, 0.0
        else:
            self.bdf_a0, self.bdf_a1, self.bdf_a2 = 1.5, -2.0, 0.5
        if self.hemo is None:
            return
        torch = self._torch
        torch.mul(self.d_un, -self.bdf_a1, out=self.d_uh)
        if self.bdf_a2 != 0.0:
            self.d_uh.add_(self.d_un2, alpha=-self.bdf_a2)
        self.hemo.set_time_scheme(1.0, self.bdf_a0, self.d_uh)

    def _after_step(self, device=False):
        # u_prev2 <- u_prev (= u^n of this step) for the next step (:323-325)
        if not device:
            self.u_prev2.x.array[:] = self.u_prev.x.array[:]
        self.d_un2.copy_(self.d_un)
        self.step_count += 1
