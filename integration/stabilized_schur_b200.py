"""Drop-in solver module for the reference tree (JuanJoZP/cfd-hemodynamic, DOLFINx v0.9 present).

Copy (or symlink) this file to `src/solvers/stabilized_schur_b200.py` of the reference, put this repository on
PYTHONPATH, and run

    python main.py simulate --simulation lid_driven2D --solver stabilized_schur_b200 --T 1.0 --dt 0.01 --name run

The reference loads `import_module(f"src.solvers.{name}").Solver` (src/scenario.py:63-78) and only touches the
`SolverBase` surface (src/solverBase.py:25-195): this `Solver` keeps that surface — it IS a `SolverBase`, its
`V / Q / u_sol / p_sol / u_prev / p_prev / u_residual / p_residual` are live DOLFINx objects the scenario's writers
and post-processing forms read (src/scenario.py:208-223,258-307) — and hands the per-timestep hot path to
libhemo_sm100.so through `cfd_hemodynamic_b200`.

What is taken from the live DOLFINx objects, so that dofmaps and the sparsity pattern are
bit-exact by construction:

    V.dofmap.list, V.dofmap.index_map_bs        cell -> block dof table (the library's "cells")
    V.tabulate_dof_coordinates()                coordinates in dof order (the library's "x")
    Q.dofmap.list                               checked / mapped onto V's numbering
    mesh.h(tdim, cells), mesh.topology.cell_name()
    compute_integration_domains(exterior_facet, ...)  (cell, local facet) pairs of each tagged ds measure
    bc.getBC(V).dof_indices(), .g.x.array, .update()  Dirichlet objects in list order

Variants: `make_solver("stabilized_schur" | "stabilized_schur_backflow" | "stabilized_schur_pressure_backflow" |
"stabilized_schur_velocity_vascular_backflow" | "stabilized_schur_bdf2")` returns the class; the module-level
`Solver` is the plain one.  A one-line module `Solver = make_solver("stabilized_schur_pressure_backflow")` gives the
hemodynamic variant under another `--solver` name.

Serial (one rank) only: under `mpirun -n N` DOLFINx has already partitioned the mesh; the multi-GPU path of this
repository partitions itself (cfd_hemodynamic_b200/distributed_solver.py).

The module is importable without DOLFINx (tests/test_integration_adapter.py drives it with duck-typed stand-ins); the
two DOLFINx functions it needs are looked up lazily and can be injected through `DOLFINX_HOOKS`.
"""
from __future__ import annotations

import importlib
from typing import Callable

import numpy as np

try:                                        # reference tree
    from src.solverBase import SolverBase as _ReferenceSolverBase
except Exception:                           # stand-alone use / tests: a minimal base with the same constructor
    _ReferenceSolverBase = None

# (mesh, facets) -> (m, 2) int array of (cell, local facet); (topology) -> exterior facet indices
DOLFINX_HOOKS: dict[str, Callable | None] = {"integration_entities": None, "exterior_facets": None}


def _integration_entities(mesh, facets):
    if DOLFINX_HOOKS["integration_entities"] is not None:
        return np.asarray(DOLFINX_HOOKS["integration_entities"](mesh, facets), dtype=np.int64).reshape(-1, 2)
    from dolfinx.fem import IntegralType, compute_integration_domains
    fdim = mesh.topology.dim - 1
    mesh.topology.create_connectivity(fdim, mesh.topology.dim)
    flat = compute_integration_domains(IntegralType.exterior_facet, mesh.topology, np.asarray(facets, dtype=np.int32), fdim)
    return np.asarray(flat, dtype=np.int64).reshape(-1, 2)


def _exterior_facets(mesh):
    if DOLFINX_HOOKS["exterior_facets"] is not None:
        return np.asarray(DOLFINX_HOOKS["exterior_facets"](mesh), dtype=np.int64)
    from dolfinx.mesh import exterior_facet_indices
    fdim = mesh.topology.dim - 1
    mesh.topology.create_connectivity(fdim, mesh.topology.dim)
    return np.asarray(exterior_facet_indices(mesh.topology), dtype=np.int64)


def _value(c):
    """float / array of a dolfinx Constant or a plain number."""
    return np.asarray(getattr(c, "value", c), dtype=np.float64)


class _ArrayHolder:
    def __init__(self, array):
        self.array = array


class _BCView:
    """One DOLFINx DirichletBC seen through the interface the B200 host layer reads: `block_dofs`, `g.x.array`
    (the SAME array DOLFINx owns: the dof numbering is shared), `dof_indices()`, `update()`."""

    def __init__(self, dbc, bs: int):
        self._dbc = dbc
        dofs = np.asarray(dbc.dof_indices()[0], dtype=np.int64)       # unrolled
        self.block_dofs = np.unique(dofs // bs)
        self._dofs = dofs
        g = dbc.g if hasattr(dbc, "g") else dbc._cpp_object.value
        self.g = type("G", (), {})()
        self.g.x = _ArrayHolder(g.x.array)

    def dof_indices(self):
        return self._dofs, len(self._dofs)

    def update(self):
        if hasattr(self._dbc, "update"):
            self._dbc.update()


class _BCAdapter:
    """Stands where the host layer expects a `BoundaryCondition`: getBC(space) returns the prepared view."""

    def __init__(self, view):
        self._view = view

    def getBC(self, _space):
        return self._view


class _TagsView:
    """facet MeshTags re-expressed in the facet numbering of the shim mesh."""

    def __init__(self, by_value: dict):
        self._by_value = by_value
        self.dim = None

    def find(self, value):
        return self._by_value.get(int(value), np.zeros(0, dtype=np.int32))


def make_solver(variant: str = "stabilized_schur"):
    inner_module = importlib.import_module(f"cfd_hemodynamic_b200.src.solvers.{variant}")
    InnerSolver = inner_module.Solver
    from cfd_hemodynamic_b200.fem.mesh import Mesh as ShimMesh
    Base = _ReferenceSolverBase if _ReferenceSolverBase is not None else object

    class Solver(Base):                      # noqa: D401 — the reference's plugin class name
        B200_VARIANT = variant

        def __init__(self, mesh, dt, rho, mu, f, initial_velocity=None, **kwargs):
            if _ReferenceSolverBase is not None:
                super().__init__(mesh, dt, rho, mu, f)
                cell = mesh.topology.cell_name()
                p_grade = int(kwargs.get("p_grade", 1))
                super().initVelocitySpace("Lagrange", cell, p_grade, shape=(mesh.geometry.dim,))
                super().initPressureSpace("Lagrange", cell, p_grade)
                if initial_velocity:
                    self.u_prev.interpolate(initial_velocity)
            else:                            # duck-typed host (tests): the caller provides the spaces / Functions
                self.mesh = mesh
                self.dt, self.rho, self.mu, self.f = dt, rho, mu, f
                host = kwargs.pop("_host_objects")
                for k, v in host.items():
                    setattr(self, k, v)
                if initial_velocity:
                    self.u_prev.interpolate(initial_velocity)
            if getattr(getattr(mesh, "comm", None), "size", 1) > 1:
                raise NotImplementedError("stabilized_schur_b200: one MPI rank only — the multi-GPU path partitions the "
                                          "mesh itself (cfd_hemodynamic_b200.distributed_solver)")
            self._bs = int(self.V.dofmap.index_map_bs)
            gdim = int(mesh.geometry.dim)
            cells_v = np.ascontiguousarray(np.asarray(self.V.dofmap.list).reshape(-1, self._nodes_per_cell(mesh)), dtype=np.int32)
            cells_q = np.asarray(self.Q.dofmap.list).reshape(cells_v.shape)
            # V and Q are separate spaces with (in practice) identical numbering; map Q onto V when they differ
            self._q_of_v = None
            if not np.array_equal(cells_v, cells_q):
                q_of_v = np.empty(int(cells_v.max()) + 1, dtype=np.int64)
                q_of_v[cells_v.reshape(-1)] = cells_q.reshape(-1)
                self._q_of_v = q_of_v
            x = np.asarray(self.V.tabulate_dof_coordinates())[:, :gdim]
            n = int(cells_v.max()) + 1
            cell_name = mesh.topology.cell_name()
            self._shim = ShimMesh(np.ascontiguousarray(x[:n]), cells_v,
                                  cell_type="quadrilateral" if cell_name == "quadrilateral" else None)
            # mesh.h of the host (what the reference's forms read, stabilized_schur.py:83-88)
            E = cells_v.shape[0]
            h_host = np.asarray(mesh.h(mesh.topology.dim, np.arange(E, dtype=np.int32)), dtype=np.float64)
            self._shim.h = lambda dim, entities, _h=h_host: _h[np.asarray(entities)]
            fval = _value(f).reshape(-1)
            self.inner = InnerSolver(self._shim, float(_value(dt)), float(_value(rho)), float(_value(mu)),
                                     [float(v) for v in fval], None, **kwargs)
            self._push_prev()

        @staticmethod
        def _nodes_per_cell(mesh):
            return {"triangle": 3, "quadrilateral": 4, "tetrahedron": 4}[mesh.topology.cell_name()]

        # ---- host <-> inner array plumbing (same dof numbering on both sides) ------------------------
        def _p_in(self, arr):
            return arr if self._q_of_v is None else np.asarray(arr)[self._q_of_v]

        def _p_out(self, dst, src):
            if self._q_of_v is None:
                dst[:len(src)] = src
            else:
                dst[self._q_of_v] = src

        def _push_prev(self):
            nu = self.inner.u_prev.x.array.shape[0]
            self.inner.u_prev.x.array[:] = np.asarray(self.u_prev.x.array)[:nu]
            self.inner.p_prev.x.array[:] = self._p_in(self.p_prev.x.array)[:self.inner.p_prev.x.array.shape[0]]

        def _pull_results(self):
            i = self.inner
            self.u_sol.x.array[:i.u_sol.x.array.shape[0]] = i.u_sol.x.array
            self.u_residual.x.array[:i.u_residual.x.array.shape[0]] = i.u_residual.x.array
            self._p_out(self.p_sol.x.array, i.p_sol.x.array)
            self._p_out(self.p_residual.x.array, i.p_residual.x.array)

        # ---- plugin API ------------------------------------------------------------------------------
        def setup(self, bcu, bcp, facet_tags=None, tags=None):
            mesh = self.mesh
            views_u = [_BCAdapter(_BCView(bc.getBC(self.V), self._bs)) for bc in bcu]
            views_p = []
            for bc in bcp:
                if self._q_of_v is not None:           # would need the values re-indexed as well
                    raise NotImplementedError("pressure Dirichlet conditions with differing V / Q dof numbering")
                views_p.append(_BCAdapter(_BCView(bc.getBC(self.Q), 1)))
            shim_tags = None
            if facet_tags is not None and tags is not None:
                c2f = self._shim.topology.cell_facets
                by_value = {}
                for name, val in tags.items():
                    if val is None:
                        continue
                    facets = np.asarray(facet_tags.find(val), dtype=np.int64)
                    if facets.size == 0:
                        continue
                    pairs = _integration_entities(mesh, facets)
                    by_value[int(val)] = np.unique(c2f[pairs[:, 0], pairs[:, 1]]).astype(np.int32)
                shim_tags = _TagsView(by_value)
            self.inner.setup(views_u, views_p, facet_tags=shim_tags, tags=tags)

        def solveStep(self):
            """One time step: host u_prev / p_prev in (the host owns the time-level shift, src/scenario.py:306-307),
            u_sol / p_sol / residuals out; RuntimeError on divergence like the reference (stabilized_schur.py:332-334)."""
            self._push_prev()
            self.inner.solveStep()
            self._pull_results()

        # bookkeeping the reference's time loop prints (stabilized_schur.py:325-330)
        @property
        def its_snes(self):
            return self.inner.its_snes

        @property
        def its_ksp(self):
            return self.inner.its_ksp

    Solver.__name__ = "Solver"
    return Solver


_SOLVER_CLASS = None


def __getattr__(name):          # `module.Solver`, built lazily: building the class imports the CUDA host layer
    global _SOLVER_CLASS
    if name == "Solver":
        if _SOLVER_CLASS is None:
            _SOLVER_CLASS = make_solver("stabilized_schur")
        return _SOLVER_CLASS
    raise AttributeError(name)
