"""`BoundaryCondition` — same surface as the reference
(src/boundaryCondition.py:14-55): wraps a value Function plus a topological or
geometrical dof locator; `getBC(V)` returns a DirichletBC whose `update()`
re-interpolates the value function."""
from types import MethodType
from typing import Callable

from numpy import dtype, int32, ndarray

from ..fem.space import (
    DirichletBC,
    Function,
    FunctionSpace,
    dirichletbc,
    locate_dofs_geometrical,
    locate_dofs_topological,
)


class BoundaryCondition:
    def __init__(self, f: Function):
        self._topological = False
        self._geometrical = False
        self.f = f

    def initTopological(self, entity_dim: int, entities: ndarray[None, dtype[int32]]) -> None:
        assert not (self._topological or self._geometrical)
        self.entity_dim = entity_dim
        self.entities = entities
        self._topological = True

    def initGeometrical(self, marker: Callable) -> None:
        assert not (self._topological or self._geometrical)
        self.marker = marker
        self._geometrical = True

    def _getDofs(self, V: FunctionSpace) -> ndarray:
        assert self._topological or self._geometrical
        if self._topological:
            return locate_dofs_topological(V, self.entity_dim, self.entities)
        if self._geometrical:
            return locate_dofs_geometrical(V, self.marker)

    def getBC(self, V: FunctionSpace) -> DirichletBC:
        dofs = self._getDofs(V)
        self._f_V = Function(V)
        self._f_V.interpolate(self.f)
        bc = dirichletbc(self._f_V, dofs)

        def update(inner_self):
            # reference: full Function->Function interpolate every residual call
            # (src/boundaryCondition.py:48-49).  Only the constrained dofs are ever
            # read, so only those are refreshed when f lives on the same space.
            f = self.f
            if isinstance(f, Function) and f.x.array.shape == self._f_V.x.array.shape:
                d, _ = inner_self.dof_indices()
                self._f_V.x.array[d] = f.x.array[d]
            else:
                self._f_V.interpolate(f)

        bc.update = MethodType(update, bc)
        return bc

    def updateBCValues(self, f: Function) -> None:
        assert self._f_V, "Boundary condition values have not been initialized."
        self.f = f
        self._f_V.interpolate(f)
