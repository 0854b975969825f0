"""`SolverBase` — the solver plugin boundary, API kept verbatim from the
reference (src/solverBase.py:25-195): ctor `(mesh, dt, rho, mu, f,
initial_velocity=None)`, abstract `setup(bcu, bcp)` / `solveStep()`,
properties `V Q u_sol p_sol u_prev p_prev`, attributes `u_residual
p_residual`, `initVelocitySpace`, `initPressureSpace`, `initStressForm`,
`assemble_wss`, static `epsilon` / `sigma`.

The reference imports DOLFINx at module scope; here the same names come from
the numpy shim (`cfd_hemodynamic_b200.fem`).
"""
from abc import ABC, abstractmethod
from typing import Callable

import numpy as np

from ..fem.mesh import Mesh, exterior_facet_indices
from ..fem.space import Constant, Function, FunctionSpace, element, functionspace
from .boundaryCondition import BoundaryCondition


class SolverBase(ABC):
    @abstractmethod
    def __init__(
        self,
        mesh: Mesh,
        dt: float,
        rho: float,
        mu: float,
        f: list,
        initial_velocity: Callable[[np.ndarray], np.ndarray] = None,
    ):
        self.mesh = mesh
        self.dt = Constant(mesh, float(dt))
        self.rho = Constant(mesh, float(rho))
        self.mu = Constant(mesh, float(mu))
        self.f = Constant(mesh, f)
        self._u_sol: Function | None = None
        self._p_sol: Function | None = None
        self._u_prev: Function | None = None
        self._p_prev: Function | None = None
        self._V: FunctionSpace | None = None
        self._Q: FunctionSpace | None = None

    @property
    def u_sol(self):
        assert self._u_sol is not None, \
            "Velocity solution function is not initialized. call initVelocitySpace() first."
        return self._u_sol

    @property
    def p_sol(self):
        assert self._p_sol is not None, \
            "Pressure solution function is not initialized. call initPressureSpace() first."
        return self._p_sol

    @property
    def u_prev(self):
        assert self._u_prev is not None, \
            "Velocity solution function is not initialized. call initVelocitySpace() first."
        return self._u_prev

    @property
    def p_prev(self):
        assert self._p_prev is not None, \
            "Pressure solution function is not initialized. call initPressureSpace() first."
        return self._p_prev

    @property
    def V(self):
        assert self._V is not None, \
            "Velocity function space is not initialized. call initVelocitySpace() first."
        return self._V

    @property
    def Q(self):
        assert self._Q is not None, \
            "Pressure function space is not initialized. call initPressureSpace() first."
        return self._Q

    @abstractmethod
    def setup(self, bcu: list[BoundaryCondition], bcp: list[BoundaryCondition]) -> None:
        pass

    @abstractmethod
    def solveStep(self) -> None:
        pass

    def initVelocitySpace(self, family, cell, deegre: int, shape: tuple[int, ...] | None = None) -> None:
        """Create `self.V`, `u_sol` ("velocity"), `u_prev`, `u_residual`."""
        element_v = element(family, cell, deegre, shape=shape)
        self._V = functionspace(self.mesh, element_v)
        self._u_sol = Function(self.V)
        self._u_sol.name = "velocity"
        self._u_prev = Function(self.V)
        self.u_residual = Function(self.V)
        self.u_residual.name = "u_residual"

    def initPressureSpace(self, family, cell, deegre: int, shape: tuple[int, ...] | None = None) -> None:
        """Create `self.Q`, `p_sol` ("pressure"), `p_prev`, `p_residual`."""
        element_p = element(family, cell, deegre, shape=shape)
        self._Q = functionspace(self.mesh, element_p)
        self._p_sol = Function(self.Q)
        self._p_sol.name = "pressure"
        self._p_prev = Function(self.Q)
        self.p_residual = Function(self.Q)
        self.p_residual.name = "p_residual"

    def initStressForm(self):
        """Wall-shear-stress output vectors (reference :144-174).  The traction
        form is evaluated on the host in `assemble_wss`; it is per-step
        post-processing outside the hot path."""
        scalar = functionspace(self.mesh, element("CG", self.mesh.topology.cell_name(), 1))
        vector = functionspace(
            self.mesh, element("CG", self.mesh.topology.cell_name(), 1, shape=(self.mesh.geometry.dim,)))
        self.normal_stress = Function(scalar)
        self.normal_stress.name = "normal_stress"
        self.shear_stress = Function(vector)
        self.shear_stress.name = "shear_stress"
        topo = self.mesh.topology
        ext = exterior_facet_indices(topo)
        self._wss_pairs = topo.facet_cell_pairs(ext)

    @staticmethod
    def epsilon(u):
        raise NotImplementedError("symbolic UFL expression; the forms are built into the CUDA kernels")

    @staticmethod
    def sigma(u, p, mu):
        raise NotImplementedError("symbolic UFL expression; the forms are built into the CUDA kernels")

    def assemble_wss(self):
        """shear_stress_i = sum_facets (1/|F|) int_F phi_i (T - (T.n) n) ds with
        T = -sigma(u, p) n (reference :144-195); P1: closed form per facet."""
        try:
            pairs = self._wss_pairs
        except AttributeError:
            return
        if self.mesh.topology.cell_name() == "tetrahedron":
            return self._assemble_wss_tetrahedron(pairs)
        x = self.mesh.geometry.x[:, :2]
        cells = self.mesh.geometry.dofmap[pairs[:, 0]]
        lf = pairs[:, 1]
        X = x[cells]
        m = cells.shape[0]
        ar = np.arange(m)
        if cells.shape[1] == 4:
            return self._assemble_wss_quadrilateral(X, cells, lf)
        fv = np.array([[1, 2], [0, 2], [0, 1]])
        va, vb = fv[lf, 0], fv[lf, 1]
        t = X[ar, vb] - X[ar, va]
        length = np.linalg.norm(t, axis=1)
        nrm = np.stack([t[:, 1], -t[:, 0]], axis=1) / length[:, None]
        sgn = np.sign(np.einsum("ei,ei->e", nrm, X[ar, va] - X[ar, lf]))
        nrm *= sgn[:, None]
        J = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]], axis=2)
        inv = np.linalg.inv(J)
        ghat = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])
        dphi = np.einsum("aj,eji->eai", ghat, inv)
        U = self.u_sol.x.array.reshape(-1, 2)[cells]
        G = np.einsum("eai,eaj->eij", dphi, U)
        mu = float(self.mu.value)
        eps = 0.5 * (G + np.swapaxes(G, 1, 2))
        # T = -sigma n = -2 mu eps n + p n; the tangential part drops p n, and
        # eps is constant on a P1 cell, so (1/|F|) int_F phi_a Tt ds = Tt / 2
        T = -2.0 * mu * np.einsum("eij,ej->ei", eps, nrm)
        Tt = T - np.einsum("ei,ei->e", T, nrm)[:, None] * nrm
        out = self.shear_stress.x.array.reshape(-1, 2)
        out[:] = 0.0
        np.add.at(out, cells[ar, va], 0.5 * Tt)
        np.add.at(out, cells[ar, vb], 0.5 * Tt)

    def _assemble_wss_tetrahedron(self, pairs):
        """The same traction form on P1 tetrahedra: eps(u) is constant on the cell, the outward normal
        of local facet lf is -grad(phi_lf) / |grad(phi_lf)|, and (1/|F|) int_F phi_a ds = 1/3 on the
        three vertices of a triangular facet."""
        cells = self.mesh.geometry.dofmap[pairs[:, 0]]
        lf = pairs[:, 1]
        X = self.mesh.geometry.x[cells]                        # (m, 4, 3)
        ar = np.arange(cells.shape[0])
        J = np.stack([X[:, j + 1] - X[:, 0] for j in range(3)], axis=2)
        ghat = np.vstack([-np.ones((1, 3)), np.eye(3)])
        dphi = np.einsum("aj,eji->eai", ghat, np.linalg.inv(J))
        nrm = -dphi[ar, lf]
        nrm /= np.linalg.norm(nrm, axis=1)[:, None]
        U = self.u_sol.x.array.reshape(-1, 3)[cells]
        G = np.einsum("eai,eaj->eij", dphi, U)
        eps = 0.5 * (G + np.swapaxes(G, 1, 2))
        T = -2.0 * float(self.mu.value) * np.einsum("eij,ej->ei", eps, nrm)
        Tt = T - np.einsum("ei,ei->e", T, nrm)[:, None] * nrm
        out = self.shear_stress.x.array.reshape(-1, 3)
        out[:] = 0.0
        for a in range(4):
            on = lf != a                                        # vertex a belongs to the facet opposite lf
            np.add.at(out, cells[on, a], Tt[on] / 3.0)

    def _assemble_wss_quadrilateral(self, X, cells, lf):
        """Same traction form on Q1 quadrilaterals: grad(u) varies along the facet, so
        (1/|F|) int_F phi_a Tt ds is integrated with 3 Gauss points per facet (the trace of
        phi_a is linear, eps(u) rational on non-affine cells)."""
        from ..fem.mesh import QUAD_FACETS
        from ..fem.quadrature import interval_gauss
        m = cells.shape[0]
        ar = np.arange(m)
        fv = np.array(QUAD_FACETS)
        va, vb = fv[lf, 0], fv[lf, 1]
        t = X[ar, vb] - X[ar, va]
        nrm = np.stack([t[:, 1], -t[:, 0]], axis=1) / np.linalg.norm(t, axis=1)[:, None]
        mid = 0.5 * (X[ar, va] + X[ar, vb])
        nrm *= np.sign(np.einsum("ei,ei->e", nrm, mid - X.mean(axis=1)))[:, None]
        U = self.u_sol.x.array.reshape(-1, 2)[cells]
        mu = float(self.mu.value)
        out = self.shear_stress.x.array.reshape(-1, 2)
        out[:] = 0.0
        for s_q, w_q in zip(*interval_gauss(3)):
            xi = np.where((lf == 0) | (lf == 3), s_q, np.where(lf == 1, 0.0, 1.0))
            eta = np.where(lf == 0, 0.0, np.where(lf == 3, 1.0, s_q))
            dref = np.stack([np.stack([-(1 - eta), -(1 - xi)], 1), np.stack([(1 - eta), -xi], 1),
                             np.stack([-eta, (1 - xi)], 1), np.stack([eta, xi], 1)], axis=1)      # (m, 4, 2)
            Jm = np.einsum("eai,eaj->eij", X, dref)
            dphi = np.einsum("eaj,eji->eai", dref, np.linalg.inv(Jm))
            G = np.einsum("eai,eaj->eij", dphi, U)
            eps = 0.5 * (G + np.swapaxes(G, 1, 2))
            T = -2.0 * mu * np.einsum("eij,ej->ei", eps, nrm)
            Tt = T - np.einsum("ei,ei->e", T, nrm)[:, None] * nrm
            np.add.at(out, cells[ar, va], (w_q * (1.0 - s_q)) * Tt)
            np.add.at(out, cells[ar, vb], (w_q * s_q) * Tt)
