"""Ethier–Steinman flow on the unit cube (reference src/scenarios/taylor_green.py:12-134): tetrahedral
`create_unit_cube(n, n, n)` (reference: n = 32), the exact velocity *and* pressure imposed on the
whole boundary and refreshed every step, initial velocity = exact solution at t = 0; the time loop
logs the relative L2 error against the exact velocity (`has_exact_solution`)."""
import numpy as np

from ...fem.mesh import create_unit_cube, exterior_facet_indices
from ...fem.space import Function
from ..boundaryCondition import BoundaryCondition
from ..scenario import Scenario


class TaylorGreenSimulation(Scenario):
    def __init__(self, solver_name, dt, T, f: tuple[float, float, float] = (0, 0, 0), *, rho=1, mu=1 / 50, n=32,
                 **solver_kwargs):
        self._mesh = None
        self._bcu = None
        self._bcp = None
        self._boundary_facets = None
        self.n = int(n)
        super().__init__(solver_name, "taylor_green", rho, mu, dt, T, f, **solver_kwargs)
        self._u_bc = Function(self.solver.V)
        self._p_bc = Function(self.solver.Q)
        self._u_bc.interpolate(self.exact_velocity(0))
        self._p_bc.interpolate(self.exact_pressure(0))
        self.setup()

    @property
    def mesh(self):
        if not self._mesh:
            self._mesh = create_unit_cube(None, self.n, self.n, self.n)
            self._mesh.topology.create_connectivity(self._mesh.topology.dim - 1, self._mesh.topology.dim)
            self._boundary_facets = exterior_facet_indices(self._mesh.topology)
        return self._mesh

    @property
    def bcu(self):
        if not self._bcu:
            bcu = BoundaryCondition(self._u_bc)
            bcu.initTopological(self.mesh.topology.dim - 1, self._boundary_facets)
            self._bcu = [bcu]
        return self._bcu

    @property
    def bcp(self):
        if not self._bcp:
            bcp = BoundaryCondition(self._p_bc)
            bcp.initTopological(self.mesh.topology.dim - 1, self._boundary_facets)
            self._bcp = [bcp]
        return self._bcp

    def initial_velocity(self, x):
        return self.exact_velocity(0)(x)

    def update_boundary_conditions(self, t):
        """Boundary data at time t (the reference does this in the after-step callback, :66-71)."""
        self._u_bc.interpolate(self.exact_velocity(t))
        self._p_bc.interpolate(self.exact_pressure(t))

    def solve(self, output_folder, afterStepCallback=None):
        def update(t):
            self.update_boundary_conditions(t)
            if afterStepCallback:
                afterStepCallback(t)

        return super().solve(output_folder, update)

    def exact_velocity(self, t):
        a, d = np.pi / 4, np.pi / 2

        def velocity(x):
            X, Y, Z = x[0], x[1], x[2]
            e = np.exp(-d * d * t)
            return np.vstack((
                -a * (np.exp(a * X) * np.sin(a * Y + d * Z) + np.exp(a * Z) * np.cos(a * X + d * Y)) * e,
                -a * (np.exp(a * Y) * np.sin(a * Z + d * X) + np.exp(a * X) * np.cos(a * Y + d * Z)) * e,
                -a * (np.exp(a * Z) * np.sin(a * X + d * Y) + np.exp(a * Y) * np.cos(a * Z + d * X)) * e))

        return velocity

    def exact_pressure(self, t):
        a, d = np.pi / 4, np.pi / 2

        def pressure(x):
            X, Y, Z = x[0], x[1], x[2]
            return (-a * a / 2 * (np.exp(2 * a * X) + np.exp(2 * a * Y) + np.exp(2 * a * Z)
                                  + 2 * np.sin(a * X + d * Y) * np.cos(a * Z + d * X) * np.exp(a * (Y + Z))
                                  + 2 * np.sin(a * Y + d * Z) * np.cos(a * X + d * Y) * np.exp(a * (Z + X))
                                  + 2 * np.sin(a * Z + d * X) * np.cos(a * Y + d * Z) * np.exp(a * (X + Y)))
                    * np.exp(-2 * d * d * t))

        return pressure
