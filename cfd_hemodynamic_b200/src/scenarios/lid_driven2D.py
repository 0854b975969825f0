"""Lid-driven cavity (reference src/scenarios/lid_driven2D.py:14-75):
`create_unit_square(nx, nx)`, no-slip walls, lid u=(1,0) on the open top
interval, no pressure BC (constant-pressure null space active)."""
import numpy as np

from ...fem.mesh import create_unit_square, locate_entities_boundary
from ...fem.space import Function
from ..boundaryCondition import BoundaryCondition
from ..scenario import Scenario


class LidDriven2DSimulation(Scenario):
    def __init__(self, solver_name, dt, T, f: tuple[float, float] = (0, 0), *, rho=1, mu=1, nx=50,
                 cell_type="triangle", **solver_kwargs):
        self._mesh = None
        self._bcu = None
        self._bcp = None
        self.Re = str(int(1 / mu)) if mu <= 1 else "0"
        self.nx = int(nx)
        self.cell_type = cell_type      # "quadrilateral": create_unit_square(..., CellType.quadrilateral)
        super().__init__(solver_name, "lid_driven2D", rho, mu, dt, T, f, **solver_kwargs)
        self.setup()

    @property
    def mesh(self):
        if not self._mesh:
            self._mesh = create_unit_square(None, self.nx, self.nx, cell_type=self.cell_type)
        return self._mesh

    @property
    def bcu(self):
        if not self._bcu:
            u_noslip = Function(self.solver.V)
            u_noslip.x.array[:] = 0
            fdim = self.mesh.topology.dim - 1
            walls_facets = locate_entities_boundary(self.mesh, fdim, self.walls)
            bc_noslip = BoundaryCondition(u_noslip)
            bc_noslip.initTopological(fdim, walls_facets)
            u_lid = Function(self.solver.V)
            u_lid.interpolate(lambda x: np.vstack((np.ones(x.shape[1]), np.zeros(x.shape[1]))))
            lid_facets = locate_entities_boundary(self.mesh, fdim, self.lid)
            bc_lid = BoundaryCondition(u_lid)
            bc_lid.initTopological(fdim, lid_facets)
            self._bcu = [bc_noslip, bc_lid]
        return self._bcu

    @property
    def bcp(self):
        if not self._bcp:
            self._bcp = []
        return self._bcp

    def initial_velocity(self, x):
        return np.zeros((self.mesh.geometry.dim, x.shape[1]), dtype=np.float64)

    @staticmethod
    def lid(x):
        return np.isclose(x[1], 1.0) & (x[0] > 1e-10) & (x[0] < 1.0 - 1e-10)

    @staticmethod
    def walls(x):
        return np.logical_or.reduce((np.isclose(x[0], 0), np.isclose(x[0], 1), np.isclose(x[1], 0)))
