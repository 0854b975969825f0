"""`--solver stabilized_schur_adaptive` on B200 (reference src/solvers/stabilized_schur_adaptive.py):
the stabilized_schur formulation (:62-125, same form) whose `solveStep` ramps the time step
linearly from 1e-4 to the target dt over the first 10 calls (:376-393) and, when a solve
diverges, retries once with 0.1*dt from the previous state (:396-426).

`dt` is a Constant of the form (time derivative and tau_supg2 = dt/2), so changing it only needs
`hemo_set_params` and the coefficient of the Schur approximation; no table is rebuilt.
"""
from typing import Callable

import numpy as np

from ._stabilized_common import StabilizedSchurB200


class Solver(StabilizedSchurB200):
    MAX_ITER = 20
    variant = "schur"
    RAMP_STEPS = 10
    MIN_DT = 1e-4

    def __init__(self, mesh, dt: float, rho: float, mu: float, f: list,
                 initial_velocity: Callable[[np.ndarray], np.ndarray] = None, **kwargs):
        super().__init__(mesh, dt, rho, mu, f, initial_velocity, **kwargs)
        self.step_count_adapt = 0
        self.target_dt = float(self.dt.value)          # :377-379

    def _set_dt(self, dt: float):
        self.dt.value = float(dt)
        if self.hemo is None:
            return
        fval = np.asarray(self.f.value, dtype=np.float64).reshape(-1)
        self.hemo.set_params(float(dt), float(self.rho.value), float(self.mu.value), fval[:2],
                             float(np.finfo(np.float64).resolution))
        lin = self.linear
        if lin is not None and lin.schur_mode == "laplace" and "schur_lap_coef" not in self._pc_kw:
            lin.opts["schur_lap_coef"] = 2.0 * float(self.rho.value) / float(dt)
            self.hemo.set_solver_opts(**lin.opts)

    def solveStep(self):
        self.step_count_adapt += 1
        if self.step_count_adapt <= self.RAMP_STEPS:
            progress = self.step_count_adapt / self.RAMP_STEPS
            new_dt = self.MIN_DT + (self.target_dt - self.MIN_DT) * progress
            self._set_dt(new_dt)
            print(f"[INFO] Adaptive DT Ramping: step {self.step_count_adapt}, dt={new_dt}")
        else:
            self._set_dt(self.target_dt)
        try:
            super().solveStep()
        except RuntimeError:
            print("[WARN] Diverged. Retrying with 0.1*dt")
            old_dt = float(self.dt.value)
            self._set_dt(0.1 * old_dt)
            n = self.n
            # reset the guess to the previous state (:402-412)
            self.d_x[:2 * n].copy_(self._pin["u_prev"], non_blocking=True)
            self.d_x[2 * n:].copy_(self._pin["p_prev"], non_blocking=True)
            try:
                super().solveStep()                       # if this fails, let it raise (:416)
            finally:
                self._set_dt(old_dt)
