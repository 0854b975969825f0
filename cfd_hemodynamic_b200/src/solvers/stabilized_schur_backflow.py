"""`--solver stabilized_schur_backflow` on B200 (reference
src/solvers/stabilized_schur_backflow.py): the stabilized_schur core without
the all-facet boundary term (:103-108, "do-nothing") plus outlet backflow
stabilization -beta_b rho (u_n.n)_- (u_m.v) added in `setup` (:158-176)."""
from typing import Callable

import numpy as np

from ...fem import discretization as D
from ._stabilized_common import SET_OUTLET
from ._stabilized_tet import StabilizedSchurTetB200 as StabilizedSchurB200      # triangles, quadrilaterals and tetrahedra


class Solver(StabilizedSchurB200):
    MAX_ITER = 20
    variant = "backflow"

    def __init__(self, mesh, dt: float, rho: float, mu: float, f: list,
                 initial_velocity: Callable[[np.ndarray], np.ndarray] = None,
                 v_max: float = None, p_grade: int = 1, beta_backflow: float = 0.2, **kwargs):
        if v_max is None:
            raise ValueError("v_max is required for stabilized_schur_backflow. "
                             "Pass it via CLI: --v_max <value>")
        self.v_max = float(v_max)
        self.beta_backflow = float(beta_backflow)
        super().__init__(mesh, dt, rho, mu, f, initial_velocity, p_grade=p_grade, **kwargs)

    def _facet_setup(self, facet_tags, tags):
        # `self.F -= ...` runs on every setup() call → multiplicity = setup count
        self._register_facets(SET_OUTLET, facet_tags.find(tags["outlet"]),
                              a_b=float(self._setup_count), beta_b=self.beta_backflow)
