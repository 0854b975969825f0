"""Stenosis with Dirichlet inlet velocity (reference
src/scenarios/stenosis_mesh_variable.py:30-450), the natural scenario for
`stabilized_schur_backflow`: parabolic inlet of peak `v_max`, no-slip walls,
do-nothing outlet with backflow stabilisation.  Plain `stabilized_schur`
ignores the pressure kwargs through **kwargs (stabilized_schur.py:51);
`stabilized_schur_velocity_vascular_backflow` additionally takes
`R_resistance` / `alpha_damping` for its resistance outlet."""
import numpy as np

from ...fem import generators
from ...fem.space import Function
from ..boundaryCondition import BoundaryCondition
from ..scenario import Scenario

_MMHG = 133.322


class StenosisMeshVariableSimulation(Scenario):
    fluid_marker = 1
    inlet_marker = 2
    outlet_marker = 3
    wall_marker = 4

    def __init__(self, solver_name, dt, T, f: tuple[float, float] = (0, 0), grade="severe", v_max: float = None,
                 beta_backflow: float = None, *, rho: float = 1.060e-3, mu: float = 3.5e-3,
                 n_elements_radial: int = 10, **kwargs):
        self._mesh = None
        self._ft = None
        self._bcu = None
        self._bcp = None
        self._v_max = v_max
        self.grade = grade
        passthrough = {k: kwargs.pop(k) for k in list(kwargs)
                       if k.startswith(("snes_", "ksp_", "amg_", "cheb_", "schur_", "pc_", "strength_", "smooth_")) or k in ("verbose", "device", "host_only", "quadrature")}
        self.mesh_options = kwargs.copy()
        self.mesh_options.setdefault("res", 2.0 * 1.57 / (2 * int(n_elements_radial)))
        # convection-dominated channel flow: the reference's SELFP matrix is the better Schur
        # approximation here (13 vs 48 outer iterations with exact sub-solves)
        passthrough.setdefault("schur_mode", "selfp")
        solver_kwargs = dict(passthrough)
        if beta_backflow is not None:
            solver_kwargs["beta_backflow"] = float(beta_backflow)
        if v_max is not None:
            solver_kwargs["v_max"] = float(v_max)
        for k in ("R_resistance", "alpha_damping"):
            if k in self.mesh_options:
                solver_kwargs[k] = float(self.mesh_options.pop(k))
        super().__init__(solver_name, "stenosis_mesh_variable", rho, mu, dt, T, f, **solver_kwargs)
        self.mesh.topology.create_connectivity(self.mesh.topology.dim - 1, self.mesh.topology.dim)
        self.setup()

    @property
    def mesh(self):
        if not self._mesh:
            self._mesh, self._ft = generators.stenosis_structured(self.grade, **self.mesh_options)
        return self._mesh

    def inlet_profile(self, x):
        R_in = self.mesh.mesh_options["R_in"]
        values = np.zeros((2, x.shape[1]), dtype=np.float64)
        r = x[1] - R_in
        values[0] = float(self._v_max or 0.0) * np.maximum(1.0 - (r / R_in) ** 2, 0.0)
        return values

    @property
    def bcu(self):
        if not self._bcu:
            fdim = self.mesh.topology.dim - 1
            u_nonslip = Function(self.solver.V)
            u_nonslip.x.array[:] = 0
            bcu_walls = BoundaryCondition(u_nonslip)
            bcu_walls.initTopological(fdim, self._ft.find(self.wall_marker))
            self._bcu = [bcu_walls]
            if self._v_max is not None:
                u_in = Function(self.solver.V)
                u_in.interpolate(self.inlet_profile)
                bcu_in = BoundaryCondition(u_in)
                bcu_in.initTopological(fdim, self._ft.find(self.inlet_marker))
                self._bcu = [bcu_in, bcu_walls]      # walls last: corner dofs stay no-slip
        return self._bcu

    @property
    def bcp(self):
        if not self._bcp:
            self._bcp = []
        return self._bcp

    def initial_velocity(self, x):
        return np.zeros((self.mesh.geometry.dim, x.shape[1]), dtype=np.float64)
