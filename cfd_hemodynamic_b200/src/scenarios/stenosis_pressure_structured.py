"""Pressure-driven stenosis on the structured transfinite mesh (reference
src/scenarios/stenosis_pressure_structured.py:30-393): weak inlet pressure +
Nitsche, resistance outlet, backflow stabilisation, no-slip walls, no Dirichlet
pressure.  Units mm-g-s; 2-D pressures use _MMHG_2D = 133.322/2 (:23-26).
The reference mesh is recombined to quadrilaterals (:379-386), which is the
default here too (`cell_type="quadrilateral"`, Q1-Q1); `cell_type="triangle"`
splits each quad into two P1 triangles."""
import numpy as np

from ...fem import generators
from ...fem.space import Function
from ..boundaryCondition import BoundaryCondition
from ..scenario import Scenario

_MMHG = 133.322
_MMHG_2D = _MMHG * 0.5


class StenosisPressureStructuredSimulation(Scenario):
    fluid_marker = 1
    inlet_marker = 2
    outlet_marker = 3
    wall_marker = 4
    scenario_name = "stenosis_pressure_structured"
    pressure_unit = _MMHG_2D
    default_cell_type = "quadrilateral"     # setRecombine, reference :384

    def __init__(self, solver_name, dt, T, f: tuple[float, float] = (0, 0), grade="severe", p_inlet: float = 80.0,
                 R_resistance: float = None, v_max: float = None, *, rho: float = 1.060e-3, mu: float = 3.5e-3,
                 **kwargs):
        self._mesh = None
        self._ft = None
        solver_keys = ("p_grade", "beta_nitsche", "beta_backflow", "alpha_damping")
        defaults = dict(p_grade=1, beta_nitsche=100.0, beta_backflow=0.2, alpha_damping=0.75)
        solver_kwargs = {k: kwargs.pop(k, defaults[k]) for k in solver_keys}
        # tolerances / preconditioner options are forwarded to the solver, the rest is mesh control
        passthrough = {k: kwargs.pop(k) for k in list(kwargs)
                       if k.startswith(("snes_", "ksp_", "amg_", "cheb_", "schur_", "pc_", "strength_", "smooth_")) or k in ("verbose", "device", "host_only",
                                                                                             "quadrature")}
        self.mesh_options = kwargs.copy()
        self.mesh_options.setdefault("cell_type", self.default_cell_type)
        self.grade = grade
        self._bcu = None
        self._bcp = None
        self._v_max = v_max
        if R_resistance is None:
            raise ValueError("R_resistance is required for pressure-driven inlet. "
                             "Pass it via CLI: --R_resistance <value>")
        solver_kwargs.update(p_inlet=float(p_inlet) * self.pressure_unit, R_resistance=float(R_resistance))
        # convection-dominated channel flow: the reference's SELFP matrix is the better Schur
        # approximation here (13 vs 48 outer iterations with exact sub-solves)
        passthrough.setdefault("schur_mode", "selfp")
        solver_kwargs.update(passthrough)
        super().__init__(solver_name, self.scenario_name, rho, mu, dt, T, list(f), **solver_kwargs)
        self.mesh.topology.create_connectivity(self.mesh.topology.dim - 1, self.mesh.topology.dim)
        self.setup()

    @property
    def mesh(self):
        if not self._mesh:
            self._mesh, self._ft = generators.stenosis_structured(self.grade, **self.mesh_options)
        return self._mesh

    @property
    def bcu(self):
        if not self._bcu:
            fdim = self.mesh.topology.dim - 1
            u_nonslip = Function(self.solver.V)
            u_nonslip.x.array[:] = 0
            bcu_walls = BoundaryCondition(u_nonslip)
            bcu_walls.initTopological(fdim, self._ft.find(self.wall_marker))
            self._bcu = [bcu_walls]
        return self._bcu

    @property
    def bcp(self):
        if not self._bcp:
            self._bcp = []
        return self._bcp

    def initial_velocity(self, x):
        if self._v_max is None:
            return np.zeros((self.mesh.geometry.dim, x.shape[1]), dtype=np.float64)
        o = self.mesh.mesh_options
        R_in, R_out, L = o["R_in"], o["R_out"], o["L"]
        x_sten, severity, slope = o["x_position_stenosis"], o["severity"], o["slope"]
        v_max = float(self._v_max)
        R_taper = R_in + (R_out - R_in) * (x[0] / L)
        r_taper_mid = R_in + (R_out - R_in) * (x_sten / L)
        h_sten = severity * r_taper_mid
        dist_x = h_sten / slope if slope > 0 else L / 4
        dist_x = max(dist_x, L * 0.05)
        dist_x = min(dist_x, min(x_sten, L - x_sten) * 0.95)
        dx_abs = np.abs(x[0] - x_sten)
        bump = np.where(dx_abs < dist_x, h_sten * 0.5 * (1.0 + np.cos(np.pi * dx_abs / dist_x)), 0.0)
        R_local = np.maximum(R_taper - bump, 1e-6)
        v_max_local = v_max * R_in / R_local
        r = x[1] - R_in
        values = np.zeros((self.mesh.geometry.dim, x.shape[1]), dtype=np.float64)
        values[0] = np.maximum(v_max_local * (1.0 - (r / R_local) ** 2), 0.0)
        return values

    def _probe_nodes(self):
        o = self.mesh.mesh_options
        x = self.mesh.geometry.x[:, :2]
        i_in = int(np.argmin(np.hypot(x[:, 0] - 0.0, x[:, 1] - o["R_in"])))
        i_out = int(np.argmin(np.hypot(x[:, 0] - o["L"], x[:, 1] - o["R_in"])))
        return i_in, i_out

    def ffr(self):
        """p at (0, R_in) and (L, R_in) and their ratio (reference :344-390)."""
        p = self.solver.p_sol.x.array
        i_in, i_out = self._probe_nodes()
        return p[i_in], p[i_out], (p[i_out] / p[i_in] if p[i_in] != 0 else float("nan"))

    def ffr_device(self):
        """The same two probes read from the device state (two scalars cross PCIe, not the field)."""
        s = self.solver
        i_in, i_out = self._probe_nodes()
        idx = s._torch.tensor([2 * s.n + i_in, 2 * s.n + i_out], device=s.hemo.device)
        p_in, p_out = (float(v) for v in s.d_x.index_select(0, idx).cpu())
        return p_in, p_out, (p_out / p_in if p_in != 0 else float("nan"))
