"""Function spaces, functions and Dirichlet conditions with the DOLFINx names
that `SolverBase`, the scenarios and `BoundaryCondition` touch.

Reference call sites: src/solverBase.py:104-142 (`functionspace`, `Function`),
src/boundaryCondition.py:33-52 (`locate_dofs_topological`, `dirichletbc`,
`Function.interpolate`), src/scenario.py:151-159 (`dofmap.index_map.size_global`,
`index_map_bs`), src/scenario.py:306-307 (`Function.x.array`).

Lagrange P1 (one dof per vertex, dof index == vertex index) on every cell type and P2 on triangles
(vertex dofs, then one dof per edge at its mid-point; cell order: vertices 0, 1, 2, then the edges
opposite them — the Basix layout).  A blocked space shares the scalar dofmap and
stores values interleaved (`bs*node + comp`), like DOLFINx.
"""
from __future__ import annotations

from typing import Callable

import numpy as np

from .mesh import Mesh, _IndexMap

__all__ = [
    "FunctionSpace", "Function", "Constant", "DirichletBC", "functionspace",
    "dirichletbc", "locate_dofs_topological", "locate_dofs_geometrical",
]


class _DofMap:
    def __init__(self, cell_dofs: np.ndarray, ndofs: int, bs: int):
        self.list = cell_dofs                # (E, ndofs_cell) int32 block indices
        self.index_map = _IndexMap(ndofs)
        self.index_map_bs = bs
        self.bs = bs

    def cell_dofs(self, c: int) -> np.ndarray:
        return self.list[c]


class FunctionSpace:
    def __init__(self, mesh: Mesh, family: str = "Lagrange", degree: int = 1,
                 shape: tuple | None = None):
        if degree not in (1, 2) or (degree == 2 and mesh.topology.cell_name() != "triangle"):
            raise NotImplementedError("the standalone host shim generates P1 spaces and P2 spaces on triangles")
        self.mesh = mesh
        self.family = family
        self.degree = degree
        self.shape = tuple(shape) if shape else ()
        bs = int(np.prod(self.shape)) if self.shape else 1
        if degree == 1:
            self._x = mesh.geometry.x
            self.dofmap = _DofMap(mesh.geometry.dofmap, mesh.geometry.x.shape[0], bs)
        else:
            from .discretization import p2_nodes
            x2, cells6 = p2_nodes(mesh)
            self._x = np.hstack([x2, np.zeros((x2.shape[0], 1))])
            self.dofmap = _DofMap(cells6, x2.shape[0], bs)
        self.value_size = bs

    @property
    def num_nodes(self) -> int:
        return self.dofmap.index_map.size_local

    def tabulate_dof_coordinates(self) -> np.ndarray:
        return self._x


def functionspace(mesh: Mesh, element) -> FunctionSpace:
    """Accepts the tuple form `("Lagrange", 1[, shape])` / `("DG", 0)` or an
    `_Element` built by `fem.space.element`."""
    if isinstance(element, _Element):
        return FunctionSpace(mesh, element.family, element.degree, element.shape)
    family, degree = element[0], element[1]
    shape = element[2] if len(element) > 2 else None
    return FunctionSpace(mesh, family, degree, shape)


class _Element:
    def __init__(self, family, cell, degree, shape=None):
        self.family, self.cell, self.degree, self.shape = family, cell, degree, shape


def element(family, cell, degree, shape=None) -> _Element:
    """`basix.ufl.element` stand-in (reference src/solverBase.py:116,136)."""
    return _Element(str(family), str(cell), int(degree), shape)


class _Vector:
    """`Function.x`: `.array` plus no-op ghost updates on one rank."""

    def __init__(self, n: int):
        self.array = np.zeros(n, dtype=np.float64)

    def scatter_forward(self):
        return None


class Function:
    def __init__(self, V: FunctionSpace, name: str = "f"):
        self.function_space = V
        self.x = _Vector(V.num_nodes * V.dofmap.index_map_bs)
        self.name = name

    def interpolate(self, u: "Callable | Function") -> None:
        """P1 nodal interpolation.  Callables receive x with shape (3, n) and
        return (value_size, n) or (n,) (reference src/scenarios/dfg_1.py:174-177)."""
        V = self.function_space
        bs = V.dofmap.index_map_bs
        if isinstance(u, Function):
            assert u.x.array.shape == self.x.array.shape
            self.x.array[:] = u.x.array
            return
        vals = np.asarray(u(V.tabulate_dof_coordinates().T), dtype=np.float64)
        n = V.num_nodes
        if bs == 1:
            self.x.array[:] = vals.reshape(-1)[:n]
        else:
            vals = vals.reshape(-1, n)
            self.x.array[:] = vals[:bs].T.reshape(-1)

    def copy(self) -> "Function":
        g = Function(self.function_space, self.name)
        g.x.array[:] = self.x.array
        return g


class Constant:
    """`dolfinx.fem.Constant` surface: `.value` (settable)."""

    def __init__(self, mesh, value):
        self._mesh = mesh
        self.value = np.asarray(value, dtype=np.float64)

    def __float__(self):
        return float(self.value)


class DirichletBC:
    """`dolfinx.fem.DirichletBC` surface: `.g` (value Function),
    `.dof_indices()` → (unrolled dofs, n_owned), `.function_space`."""

    def __init__(self, g: Function, block_dofs: np.ndarray):
        self.g = g
        self.function_space = g.function_space
        self._block_dofs = np.unique(np.asarray(block_dofs, dtype=np.int32))

    @property
    def block_dofs(self) -> np.ndarray:
        return self._block_dofs

    def dof_indices(self):
        bs = self.function_space.dofmap.index_map_bs
        d = (self._block_dofs[:, None] * bs + np.arange(bs, dtype=np.int32)[None, :]).reshape(-1)
        return d.astype(np.int32), d.shape[0]


def dirichletbc(value: Function, dofs: np.ndarray, V: FunctionSpace | None = None) -> DirichletBC:
    return DirichletBC(value, dofs)


def locate_dofs_topological(V: FunctionSpace, entity_dim: int, entities) -> np.ndarray:
    """Block dof indices on the closure of the listed facets (P1: the facet
    vertices).  3P `dolfinx.fem.locate_dofs_topological`; trigger: reference
    src/boundaryCondition.py:36."""
    topo = V.mesh.topology
    assert entity_dim == topo.dim - 1, "only facet entities are used by the scenarios"
    entities = np.asarray(entities, dtype=np.int64)
    fv = topo.facet_vertices[entities]
    dofs = fv.reshape(-1)
    if getattr(V, "degree", 1) == 2:         # P2 triangles: plus the dof of the facet's own edge
        dofs = np.concatenate([dofs, V.mesh.geometry.x.shape[0] + entities])
    return np.unique(dofs).astype(np.int32)


def locate_dofs_geometrical(V: FunctionSpace, marker) -> np.ndarray:
    x = V.tabulate_dof_coordinates().T
    return np.nonzero(np.asarray(marker(x), dtype=bool))[0].astype(np.int32)
