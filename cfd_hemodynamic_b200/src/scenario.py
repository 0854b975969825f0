"""`Scenario` — problem definition + the time loop, API kept from the
reference (src/scenario.py:20-360): abstract `mesh`, `bcu`, `bcp`,
`initial_velocity`; concrete `setup()` and `solve()`.  The solver is loaded by
module name exactly like the reference does (:62-100), from this package's
`src.solvers`.  VTX/ADIOS2 output (host I/O, out of scope) is replaced by a
plain `.npy` writer with the same write-every-step call pattern.
"""
import inspect
import os
import sys
from abc import ABC, abstractmethod
from importlib import import_module
from typing import Callable

import numpy as np

from ..fem.space import Function
from .boundaryCondition import BoundaryCondition
from .solverBase import SolverBase


class NpyWriter:
    """Stand-in for `dolfinx.io.VTXWriter(comm, path, function)`."""

    def __init__(self, comm, path: str, function: Function, enabled: bool = True):
        self.path = path
        self.function = function
        self.enabled = enabled
        self.count = 0
        if enabled:
            os.makedirs(path, exist_ok=True)

    def write(self, t: float):
        if self.enabled:
            np.save(os.path.join(self.path, f"step_{self.count:06d}.npy"), self.function.x.array)
        self.count += 1

    def close(self):
        return None


def l2_norm_sq(mesh, u: Function) -> float:
    """assemble_scalar(form(inner(u, u) * dx)) for P1 functions (exact)."""
    cells = mesh.geometry.dofmap
    if mesh.topology.cell_name() == "tetrahedron":
        X = mesh.geometry.x[cells]                             # (E, 4, 3)
        det = np.abs(np.linalg.det(np.stack([X[:, j + 1] - X[:, 0] for j in range(3)], axis=2)))
        bs = u.function_space.dofmap.index_map_bs
        vals = u.x.array.reshape(-1, bs)[cells]                # (E, 4, bs)
        Mloc = (np.ones((4, 4)) + np.eye(4)) / 120.0           # int phi_a phi_b over the reference tetrahedron
        return float(np.einsum("e,ab,eak,ebk->", det, Mloc, vals, vals))
    x = mesh.geometry.x[:, :2]
    X = x[cells]
    if cells.shape[1] == 4:
        # Q1 quadrilaterals: 3 x 3 Gauss points (exact on affine cells)
        bs = u.function_space.dofmap.index_map_bs
        vals = u.x.array.reshape(-1, bs)[cells]                # (E, 4, bs)
        gp = 0.5 + 0.5 * np.array([-np.sqrt(0.6), 0.0, np.sqrt(0.6)])
        gw = np.array([5.0, 8.0, 5.0]) / 18.0
        total = 0.0
        for xi, wx in zip(gp, gw):
            for eta, wy in zip(gp, gw):
                phi = np.array([(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta])
                dref = np.array([[-(1 - eta), -(1 - xi)], [(1 - eta), -xi], [-eta, (1 - xi)], [eta, xi]])
                Jm = np.einsum("eai,aj->eij", X, dref)
                det = np.abs(Jm[:, 0, 0] * Jm[:, 1, 1] - Jm[:, 0, 1] * Jm[:, 1, 0])
                f = np.einsum("a,eak->ek", phi, vals)
                total += float(np.sum(wx * wy * det * np.einsum("ek,ek->e", f, f)))
        return total
    det = np.abs((X[:, 1, 0] - X[:, 0, 0]) * (X[:, 2, 1] - X[:, 0, 1])
                 - (X[:, 2, 0] - X[:, 0, 0]) * (X[:, 1, 1] - X[:, 0, 1]))
    bs = u.function_space.dofmap.index_map_bs
    vals = u.x.array.reshape(-1, bs)[cells]                    # (E, 3, bs)
    Mloc = (np.ones((3, 3)) + np.eye(3)) / 24.0
    return float(np.einsum("e,ab,eak,ebk->", det, Mloc, vals, vals))


class Scenario(ABC):
    @property
    @abstractmethod
    def mesh(self):
        pass

    @property
    @abstractmethod
    def bcu(self) -> list[BoundaryCondition]:
        pass

    @property
    @abstractmethod
    def bcp(self) -> list[BoundaryCondition]:
        pass

    @abstractmethod
    def initial_velocity(self, x: np.ndarray) -> np.ndarray:
        pass

    def exact_velocity(self, t):
        pass

    def __init__(self, solver_name: str, scenario_name: str, rho: float, mu: float, dt: float, T: float,
                 f: list, early_stop_tolerance: float = 1e-3, **solver_kwargs):
        self.solver_name = solver_name
        self.scenario_name = scenario_name
        self.early_stop_tolerance = early_stop_tolerance
        self.write_output = True
        pkg = __name__.rsplit(".", 1)[0]
        try:
            solver_module = import_module(f"{pkg}.solvers.{solver_name}")
        except ImportError as e:
            available = self._list_available_solvers()
            raise ImportError(
                f"Could not import solver '{solver_name}'. "
                f"Ensure src/solvers/{solver_name}.py exists and all its dependencies are available.\n"
                f"Underlying error: {e}\nAvailable solvers: {available}") from e
        if not hasattr(solver_module, "Solver"):
            raise ValueError(f"Solver module 'src/solvers/{solver_name}.py' does not define a 'Solver' class.")
        self.solverClass: type[SolverBase] = solver_module.Solver

        sig = inspect.signature(self.solverClass.__init__)
        accepted = sig.parameters
        has_var_keyword = any(p.kind == inspect.Parameter.VAR_KEYWORD for p in accepted.values())
        filtered_kwargs = (solver_kwargs if has_var_keyword
                           else {k: v for k, v in solver_kwargs.items() if k in accepted})
        try:
            self.solver = self.solverClass(self.mesh, dt, rho, mu, f,
                                           initial_velocity=self.initial_velocity, **filtered_kwargs)
        except TypeError as e:
            raise RuntimeError(
                f"Failed to instantiate solver '{solver_name}': {e}. "
                f"Check that the Solver class has the correct constructor signature.") from e
        except Exception as e:
            raise RuntimeError(
                f"Error while initializing solver '{solver_name}': {type(e).__name__}: {e}") from e
        self.T = T
        self.has_exact_solution = self.__class__.exact_velocity is not Scenario.exact_velocity
        self.dt = dt

    @staticmethod
    def _list_available_solvers():
        solvers_dir = os.path.join(os.path.dirname(__file__), "solvers")
        try:
            files = os.listdir(solvers_dir)
            solvers = [f[:-3] for f in files if f.endswith(".py") and not f.startswith("_")]
            return solvers if solvers else ["(none found)"]
        except OSError:
            return ["(could not list)"]

    @property
    def facet_tags(self):
        return getattr(self, "_ft", None)

    @property
    def tags(self) -> dict:
        return {
            "inlet": getattr(self, "inlet_marker", None),
            "outlet": getattr(self, "outlet_marker", None),
            "wall": getattr(self, "wall_marker", None),
            "obstacle": getattr(self, "obstacle_marker", None),
        }

    def setup(self):
        self.solver.setup(self.bcu, self.bcp, facet_tags=self.facet_tags, tags=self.tags)
        if self.mesh.comm.rank == 0:
            num_dofs_V = self.solver.V.dofmap.index_map.size_global * self.solver.V.dofmap.index_map_bs
            num_dofs_Q = self.solver.Q.dofmap.index_map.size_global * self.solver.Q.dofmap.index_map_bs
            total_dofs = num_dofs_V + num_dofs_Q
            print(f"DOFs: {total_dofs} (Velocity: {num_dofs_V}, Pressure: {num_dofs_Q})")
            print(f"Suggested cores: {total_dofs / 20000:.1f}")

    def solve(self, output_folder: str, afterStepCallback: Callable[[float], None] = None) -> str:
        """The reference time loop (src/scenario.py:166-331), statement by
        statement: write initial state, `while t < T: solveStep(); ...;
        u_prev <- u_sol`, early stop every 10th step, final L2 norms."""
        mesh = self.mesh
        T = self.T
        solver = self.solver
        if mesh.comm.rank == 0:
            os.makedirs(output_folder, exist_ok=True)
        mesh.comm.barrier()
        wo = self.write_output
        u_file = NpyWriter(mesh.comm, f"{output_folder}/v.bp", solver.u_sol, wo)
        p_file = NpyWriter(mesh.comm, f"{output_folder}/p.bp", solver.p_sol, wo)
        u_res_file = NpyWriter(mesh.comm, f"{output_folder}/u_residual.bp", solver.u_residual, wo)
        p_res_file = NpyWriter(mesh.comm, f"{output_folder}/p_residual.bp", solver.p_residual, wo)
        solver.initStressForm()
        wss_file = NpyWriter(mesh.comm, f"{output_folder}/wss.bp", solver.shear_stress, wo)

        t = 0.0
        solver.u_sol.interpolate(self.initial_velocity)
        solver.assemble_wss()
        for w in (u_file, p_file, u_res_file, p_res_file, wss_file):
            w.write(t)

        error_log = None
        if self.has_exact_solution:
            error_log = open(f"{output_folder}/err.txt", "w") if mesh.comm.rank == 0 else None
            u_e = Function(solver.V)
            u_e.interpolate(lambda x: self.exact_velocity(t)(x))
            error = self.compute_error(solver.u_sol, u_e, mesh)
            if error_log:
                error_log.write("t = %.3f: error = %.3g" % (t, error) + "\n")

        i = 0
        while t < T:
            solver.solveStep()
            i += 1
            t += self.dt
            if self.has_exact_solution:
                u_e.interpolate(self.exact_velocity(t))
                error = self.compute_error(u_e, solver.u_sol, mesh)
                if error_log:
                    error_log.write("t = %.3f: error = %.3g" % (t, error) + "\n")
            solver.assemble_wss()
            for w in (u_file, p_file, u_res_file, p_res_file, wss_file):
                w.write(t)
            if afterStepCallback:
                afterStepCallback(t)
            if (i + 1) % 10 == 0:
                u_sol_arr = solver.u_sol.x.array
                u_prev_arr = solver.u_prev.x.array
                u_sol_norm = mesh.comm.allreduce(np.linalg.norm(u_sol_arr, ord=np.inf))
                u_diff_norm = mesh.comm.allreduce(np.linalg.norm(u_sol_arr - u_prev_arr, ord=np.inf))
                rel_diff = (u_diff_norm / max(u_sol_norm, 1e-12)) / self.dt
                if rel_diff < self.early_stop_tolerance:
                    print(f"Early stopping at t={t:.3f}, because (||u_sol - u_prev||_inf / ||u_sol||_inf) / dt "
                          f"= {rel_diff:.20e} < {self.early_stop_tolerance}")
                    break
            solver.u_prev.x.array[:] = solver.u_sol.x.array[:]
            solver.p_prev.x.array[:] = solver.p_sol.x.array[:]

        for w in (u_file, p_file, u_res_file, p_res_file, wss_file):
            w.close()
        norm_v = np.sqrt(mesh.comm.allreduce(l2_norm_sq(mesh, solver.u_sol)))
        norm_p = np.sqrt(mesh.comm.allreduce(l2_norm_sq(mesh, solver.p_sol)))
        if mesh.comm.rank == 0:
            with open(os.path.join(output_folder, "norms.txt"), "w") as f:
                f.write(f"L2 norm of velocity: {norm_v}\n")
                f.write(f"L2 norm of pressure: {norm_p}\n")
        if error_log:
            error_log.close()
        self.steps_done = i
        return output_folder

    def solve_device(self, output_folder: str | None = None, afterStepCallback: Callable[[float], None] = None):
        """The same time loop as `solve` with the state resident on the device: `step_device`,
        wall shear stress, the early-stop test and the final L2 norms all run as kernels
        (SURVEY.md §8(f) rank 2); the fields are copied to the host Functions once, at the end.
        No per-step output files (that is what `solve` is for); `norms.txt` is written when
        `output_folder` is given.  Returns (steps, norm_v, norm_p)."""
        solver = self.solver
        if solver.hemo is None:
            raise RuntimeError("solve_device needs a device context (host_only solver)")
        solver.initStressForm()
        t = 0.0
        i = 0
        while t < self.T:
            solver.step_device(shift=False)
            i += 1
            t += self.dt
            solver.assemble_wss_device()
            if afterStepCallback:
                afterStepCallback(t)
            if (i + 1) % 10 == 0:
                u_diff_norm, u_sol_norm = solver.early_stop_norms_device()
                rel_diff = (u_diff_norm / max(u_sol_norm, 1e-12)) / self.dt
                if rel_diff < self.early_stop_tolerance:
                    print(f"Early stopping at t={t:.3f}, because (||u_sol - u_prev||_inf / ||u_sol||_inf) / dt "
                          f"= {rel_diff:.20e} < {self.early_stop_tolerance}")
                    break
            solver.shift_time_level_device()
        norm_v, norm_p = solver.l2_norms_device()
        solver.download_solution()
        if output_folder is not None and self.mesh.comm.rank == 0:
            os.makedirs(output_folder, exist_ok=True)
            with open(os.path.join(output_folder, "norms.txt"), "w") as f:
                f.write(f"L2 norm of velocity: {norm_v}\n")
                f.write(f"L2 norm of pressure: {norm_p}\n")
        self.steps_done = i
        return i, norm_v, norm_p

    @staticmethod
    def compute_error(u: Function, u_aprox: Function, mesh) -> float:
        """Relative L2 error between u and u_aprox (reference :350-360)."""
        d = Function(u.function_space)
        d.x.array[:] = u_aprox.x.array - u.x.array
        return float(np.sqrt(l2_norm_sq(mesh, d)) / np.sqrt(l2_norm_sq(mesh, u)))
