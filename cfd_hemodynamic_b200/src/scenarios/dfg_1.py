"""DFG 2D-1 benchmark, cylinder in a channel at Re = 20 (reference
src/scenarios/dfg_1.py:17-255): parabolic inlet, no-slip walls and obstacle,
p = 0 Dirichlet at the outlet, drag/lift/pressure-difference post-processing
with the reference's formulas and scale 500 (:183-253)."""
import os

import numpy as np

from ...fem import generators
from ...fem.space import Function
from ..boundaryCondition import BoundaryCondition
from ..scenario import Scenario


class DFG1Benchmark(Scenario):
    fluid_marker = 1
    inlet_marker = 2
    outlet_marker = 3
    wall_marker = 4
    obstacle_marker = 5

    def __init__(self, solver_name, dt, T, f: tuple[float, float] = (0, 0), *, rho=1, mu=1 / 1000,
                 lc_min=None, lc_max=None, **solver_kwargs):
        self._mesh = None
        self._ft = None
        self._bcu = None
        self._bcp = None
        self.mu = mu
        self.rho = rho
        self._lc = (lc_min, lc_max)
        super().__init__(solver_name, "dfg_1", rho, mu, dt, T, f, **solver_kwargs)
        self.mesh.topology.create_connectivity(self.mesh.topology.dim - 1, self.mesh.topology.dim)
        self.setup()

    @property
    def mesh(self):
        if not self._mesh:
            self._mesh, self._ft = generators.dfg_cylinder(*self._lc)
        return self._mesh

    @property
    def bcu(self):
        if not self._bcu:
            fdim = self.mesh.topology.dim - 1
            u_inlet = Function(self.solver.V)
            u_inlet.interpolate(self.inlet_velocity)
            bcu_inflow = BoundaryCondition(u_inlet)
            bcu_inflow.initTopological(fdim, self._ft.find(self.inlet_marker))
            u_nonslip = Function(self.solver.V)
            u_nonslip.x.array[:] = 0
            bcu_walls = BoundaryCondition(u_nonslip)
            bcu_walls.initTopological(fdim, self._ft.find(self.wall_marker))
            bcu_obstacle = BoundaryCondition(u_nonslip)
            bcu_obstacle.initTopological(fdim, self._ft.find(self.obstacle_marker))
            self._bcu = [bcu_inflow, bcu_obstacle, bcu_walls]
        return self._bcu

    @property
    def bcp(self):
        if not self._bcp:
            fdim = self.mesh.topology.dim - 1
            pr = Function(self.solver.Q)
            pr.x.array[:] = 0
            bc_outflow = BoundaryCondition(pr)
            bc_outflow.initTopological(fdim, self._ft.find(self.outlet_marker))
            self._bcp = [bc_outflow]
        return self._bcp

    def initial_velocity(self, x):
        return np.zeros((self.mesh.geometry.dim, x.shape[1]), dtype=np.float64)

    @staticmethod
    def inlet_velocity(x):
        values = np.zeros((2, x.shape[1]), dtype=np.float64)
        values[0] = 4 * 0.3 * x[1] * (0.41 - x[1]) / (0.41 ** 2)
        return values

    def drag_lift(self):
        """F_D, F_L of dfg_1.py:183-202 (P1: grad u constant per cell), times 500."""
        mesh = self.mesh
        pairs = mesh.topology.facet_cell_pairs(self._ft.find(self.obstacle_marker))
        x = mesh.geometry.x[:, :2]
        cells = mesh.geometry.dofmap[pairs[:, 0]]
        lf = pairs[:, 1]
        X = x[cells]
        ar = np.arange(cells.shape[0])
        fv = np.array([[1, 2], [0, 2], [0, 1]])
        va, vb = fv[lf, 0], fv[lf, 1]
        t = X[ar, vb] - X[ar, va]
        length = np.linalg.norm(t, axis=1)
        nout = np.stack([t[:, 1], -t[:, 0]], axis=1) / length[:, None]
        nout *= np.sign(np.einsum("ei,ei->e", nout, X[ar, va] - X[ar, lf]))[:, None]
        n = -nout                                      # n = -FacetNormal (:191)
        tang = np.stack([n[:, 1], -n[:, 0]], axis=1)
        J = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]], axis=2)
        dphi = np.einsum("aj,eji->eai", np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]]), np.linalg.inv(J))
        U = self.solver.u_sol.x.array.reshape(-1, 2)[cells]
        P = self.solver.p_sol.x.array[cells]
        G = np.einsum("eai,eaj->eij", dphi, U)         # d_i u_j
        dut_dn = np.einsum("ei,eij,ej->e", n, G, tang)  # grad(u_t) . n
        pbar = 0.5 * (P[ar, va] + P[ar, vb])
        FD = float(np.sum(length * (self.mu * dut_dn * n[:, 1] - pbar * n[:, 0])))
        FL = float(np.sum(-length * (self.mu * dut_dn * n[:, 0] + pbar * n[:, 1])))
        return 500 * FD, 500 * FL

    def drag_lift_consistent(self):
        """Cd, Cl from the consistent nodal forces on the cylinder (`Solver.consistent_boundary_force`), times 500."""
        nodes = np.unique(self.mesh.topology.facet_vertices[self._ft.find(self.obstacle_marker)])
        fx, fy = self.solver.consistent_boundary_force(nodes)
        return 500 * fx, 500 * fy

    def drag_lift_device(self):
        """Same integrals evaluated by `hemo_boundary_force` from the device state (no D2H of the fields)."""
        fd, fl = self.solver.boundary_force_device(self._ft.find(self.obstacle_marker))
        return 500 * fd, 500 * fl

    def pressure_difference(self):
        """p(0.15, 0.2) - p(0.25, 0.2) (dfg_1.py:213-253) by P1 interpolation."""
        mesh = self.mesh
        x = mesh.geometry.x[:, :2]
        cells = mesh.geometry.dofmap
        out = []
        for pt in ((0.15, 0.2), (0.25, 0.2)):
            X = x[cells]
            T = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]], axis=2)
            lam = np.linalg.solve(T, (np.asarray(pt) - X[:, 0])[:, :, None])[:, :, 0]
            l0 = 1 - lam.sum(axis=1)
            inside = np.nonzero((lam >= -1e-10).all(axis=1) & (l0 >= -1e-10))[0]
            c = inside[0]
            P = self.solver.p_sol.x.array[cells[c]]
            out.append(l0[c] * P[0] + lam[c, 0] * P[1] + lam[c, 1] * P[2])
        return float(out[0] - out[1])

    def solve(self, output_folder, afterStepCallback=None):
        out_path = super().solve(output_folder, afterStepCallback)
        cd, cl = self.drag_lift()
        print(f"Drag: {cd}")
        print(f"Lift: {cl}")
        with open(f"{out_path}/drag_lift.txt", "w") as f:
            f.write(f"Drag: {cd}\n")
            f.write(f"Lift: {cl}\n")
        dp = self.pressure_difference()
        print(f"Pressure difference: {dp}")
        with open(f"{out_path}/pressure_diff.txt", "w") as f:
            f.write(f"Pressure difference: {dp}\n")
        return out_path
