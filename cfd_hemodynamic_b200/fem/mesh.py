"""Host-side mesh container with the DOLFINx attribute surface the hot path reads.

The reference hands a `dolfinx.mesh.Mesh` to `Solver.__init__`
(reference: src/solvers/stabilized_schur.py:43-58) and reads
`mesh.geometry.x / .dofmap`, `mesh.topology.dim / cell_name()`, `mesh.comm`
and `mesh.h(tdim, cells)` (:83-88).  DOLFINx is not installable in this image,
so this module provides a plain-numpy object with the same attribute names;
when a real dolfinx mesh is passed to the solver the adapter in
`fem/adapter.py` pulls the same arrays from it instead.

Local facet numbering follows Basix: simplex facet i is opposite vertex i.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "Mesh", "MeshTags", "SerialComm", "meshtags", "locate_entities_boundary",
    "exterior_facet_indices", "create_unit_square", "create_rectangle",
    "create_mesh", "create_unit_cube", "create_box",
]


class SerialComm:
    """Stand-in for `mesh.comm` (mpi4py communicator) on one rank.

    The time loop only uses rank, size, barrier and scalar allreduce
    (reference: src/scenario.py:206,273-280).  Under torchrun the solver swaps
    in `parallel.TorchComm`, which implements the same calls on
    torch.distributed.
    """

    rank = 0
    size = 1

    def barrier(self):
        return None

    def Barrier(self):
        return None

    def allreduce(self, value, op=None):
        return value

    def bcast(self, value, root=0):
        return value

    def gather(self, value, root=0):
        return [value]


class _Geometry:
    def __init__(self, x: np.ndarray, dofmap: np.ndarray, gdim: int):
        self.x = x            # (n, 3) float64, z-padded like dolfinx
        self.dofmap = dofmap  # (E, nodes_per_cell) int32
        self.dim = gdim


class _IndexMap:
    def __init__(self, n: int):
        self.size_local = int(n)
        self.num_ghosts = 0
        self.size_global = int(n)
        self.local_range = (0, int(n))
        self.ghosts = np.zeros(0, dtype=np.int64)
        self.owners = np.zeros(0, dtype=np.int32)


# local facets of a tensor-ordered quadrilateral (3P Basix numbering)
QUAD_FACETS = ((0, 1), (0, 2), (1, 3), (2, 3))


class _Topology:
    """Cell/facet/vertex connectivity for simplices (triangle, tetrahedron) and
    tensor-ordered quadrilaterals."""

    def __init__(self, cells: np.ndarray, tdim: int, nverts: int, cell_type: str | None = None):
        self.dim = tdim
        self._cell_type = cell_type or {2: "triangle", 3: "tetrahedron"}[tdim]
        self._cells = cells
        self._nverts = nverts
        self._facets = None
        self._f2c_count = None
        self._c2f = None
        self._f2c_first = None
        self._f2c_local = None

    def cell_name(self) -> str:
        return self._cell_type

    # -- lazily built facet tables ------------------------------------
    def _build_facets(self):
        if self._facets is not None:
            return
        c = self._cells
        if self._cell_type == "quadrilateral":
            loc = [list(f) for f in QUAD_FACETS]
        else:
            # local facet i = all vertices except vertex i
            loc = [[j for j in range(c.shape[1]) if j != i] for i in range(c.shape[1])]
        nfv = len(loc[0])                                        # vertices per facet
        fv = np.stack([c[:, l] for l in loc], axis=1)            # (E, facets per cell, nfv)
        self._finish_facets(np.sort(fv.reshape(-1, nfv), axis=1), nfv, len(loc))

    def _finish_facets(self, fv_sorted, nfv, nv):
        """fv_sorted: (E*nv, nfv) sorted vertex tuples of every (cell, local facet)."""
        c = self._cells
        # pack the sorted vertex tuple into one int64 key (1-D unique is much
        # faster than axis=0 unique on 10^7 facets)
        n = np.int64(self._nverts)
        key = fv_sorted[:, 0].astype(np.int64)
        for k in range(1, nfv):
            key = key * n + fv_sorted[:, k]
        ukey, inv, counts = np.unique(key, return_inverse=True, return_counts=True)
        inv = inv.reshape(-1)
        uniq = np.empty((ukey.shape[0], nfv), dtype=np.int64)
        rem = ukey
        for k in range(nfv - 1, -1, -1):
            uniq[:, k] = rem % n
            rem = rem // n
        self._facets = uniq.astype(np.int32)
        self._c2f = inv.reshape(c.shape[0], nv).astype(np.int32)
        self._f2c_count = counts.astype(np.int32)
        # first incident (cell, local facet) of each facet
        order = np.argsort(inv, kind="stable")
        first = np.searchsorted(inv[order], np.arange(uniq.shape[0]))
        flat = order[first]
        self._f2c_first = (flat // nv).astype(np.int32)
        self._f2c_local = (flat % nv).astype(np.int32)

    def create_connectivity(self, d0, d1):
        self._build_facets()

    def create_entities(self, d):
        self._build_facets()

    def index_map(self, d: int) -> _IndexMap:
        if d == self.dim:
            return _IndexMap(self._cells.shape[0])
        if d == 0:
            return _IndexMap(self._nverts)
        if d == self.dim - 1:
            self._build_facets()
            return _IndexMap(self._facets.shape[0])
        raise NotImplementedError(f"index_map({d})")

    @property
    def facet_vertices(self) -> np.ndarray:
        self._build_facets()
        return self._facets

    @property
    def cell_facets(self) -> np.ndarray:
        self._build_facets()
        return self._c2f

    def facet_cell_pairs(self, facets: np.ndarray) -> np.ndarray:
        """(cell, local_facet) of the first cell attached to each facet —
        the integration entities of an exterior-facet integral (3P:
        dolfinx.fem.compute_integration_domains)."""
        self._build_facets()
        facets = np.asarray(facets, dtype=np.int64)
        return np.stack([self._f2c_first[facets], self._f2c_local[facets]],
                        axis=1).astype(np.int32)


class Mesh:
    """Simplicial or quadrilateral mesh; degree-1 geometry (geometry dofmap ==
    vertex list).  `cell_type="quadrilateral"`: 4 tensor-ordered vertices per
    cell, (0,0),(1,0),(0,1),(1,1) — what gmshio.model_to_mesh yields for the
    recombined transfinite mesh of the reference
    (src/scenarios/stenosis_pressure_structured.py:379-386)."""

    def __init__(self, x: np.ndarray, cells: np.ndarray, comm=None, cell_type: str | None = None):
        x = np.asarray(x, dtype=np.float64)
        if x.shape[1] == 2:
            x = np.hstack([x, np.zeros((x.shape[0], 1))])
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        if cell_type == "quadrilateral":
            if cells.shape[1] != 4:
                raise ValueError("quadrilateral cells need 4 vertices")
            tdim = 2
        else:
            tdim = cells.shape[1] - 1
        # all BASELINE configs have gdim == tdim
        self.geometry = _Geometry(np.ascontiguousarray(x), cells, tdim)
        self.topology = _Topology(cells, tdim, x.shape[0], cell_type)
        self.comm = comm if comm is not None else SerialComm()
        self.name = "mesh"

    @property
    def num_cells(self) -> int:
        return self.geometry.dofmap.shape[0]

    @property
    def num_vertices(self) -> int:
        return self.geometry.x.shape[0]

    def h(self, dim: int, entities: np.ndarray) -> np.ndarray:
        """Max vertex-vertex distance per cell (3P `Mesh.h`; trigger:
        reference src/solvers/stabilized_schur.py:85-88)."""
        assert dim == self.topology.dim
        c = self.geometry.dofmap[np.asarray(entities, dtype=np.int64)]
        X = self.geometry.x[c]                      # (n, nv, 3)
        nv = c.shape[1]
        h = np.zeros(c.shape[0])
        for i in range(nv):
            for j in range(i + 1, nv):
                h = np.maximum(h, np.linalg.norm(X[:, i] - X[:, j], axis=1))
        return h


class MeshTags:
    """`dolfinx.mesh.MeshTags` surface: dim, indices (sorted), values, find."""

    def __init__(self, mesh: Mesh, dim: int, indices, values):
        indices = np.asarray(indices, dtype=np.int32)
        values = np.asarray(values, dtype=np.int32)
        order = np.argsort(indices, kind="stable")
        self.mesh = mesh
        self.dim = dim
        self.indices = indices[order]
        self.values = values[order]
        self.name = "facet_tags"

    def find(self, value: int) -> np.ndarray:
        return self.indices[self.values == value]


def meshtags(mesh, dim, indices, values) -> MeshTags:
    return MeshTags(mesh, dim, indices, values)


def exterior_facet_indices(topology: _Topology) -> np.ndarray:
    topology._build_facets()
    return np.nonzero(topology._f2c_count == 1)[0].astype(np.int32)


def locate_entities_boundary(mesh: Mesh, dim: int, marker) -> np.ndarray:
    """Exterior facets whose vertices all satisfy `marker(x)` with x of shape
    (3, n) (3P `dolfinx.mesh.locate_entities_boundary`; trigger: reference
    src/scenarios/lid_driven2D.py:41,49)."""
    assert dim == mesh.topology.dim - 1
    ext = exterior_facet_indices(mesh.topology)
    fv = mesh.topology.facet_vertices[ext]
    marked_v = np.asarray(marker(mesh.geometry.x.T), dtype=bool)
    keep = np.all(marked_v[fv], axis=1)
    return ext[keep]


def create_mesh(x, cells, comm=None, cell_type: str | None = None) -> Mesh:
    return Mesh(x, cells, comm, cell_type)


def create_rectangle(p0, p1, nx: int, ny: int, diagonal: str = "right", comm=None,
                     cell_type: str = "triangle") -> Mesh:
    """Structured mesh of a rectangle.  Triangles: "right" diagonal matches
    `dolfinx.mesh.create_unit_square` default used by the lid cavity
    (reference src/scenarios/lid_driven2D.py:30).  `cell_type="quadrilateral"`
    gives the tensor-ordered Q1 grid (`CellType.quadrilateral`, reference
    src/scenarios/unit_square.py:36-38)."""
    xs = np.linspace(p0[0], p1[0], nx + 1)
    ys = np.linspace(p0[1], p1[1], ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    pts = np.stack([X.ravel(), Y.ravel()], axis=1)
    if cell_type == "quadrilateral":
        return Mesh(pts, grid_quads(nx, ny), comm, cell_type)
    cells = _split_grid(nx, ny, diagonal)
    return Mesh(pts, cells, comm)


def grid_quads(nx: int, ny: int) -> np.ndarray:
    """(nx*ny, 4) tensor-ordered quadrilaterals of an (nx+1) x (ny+1) vertex grid
    numbered row by row (x fastest)."""
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v0 = (iy * (nx + 1) + ix).ravel()
    return np.stack([v0, v0 + 1, v0 + (nx + 1), v0 + (nx + 2)], axis=1).astype(np.int32)


def _split_grid(nx: int, ny: int, diagonal: str = "right") -> np.ndarray:
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v0 = (iy * (nx + 1) + ix).ravel()
    v1 = v0 + 1
    v2 = v0 + (nx + 1)
    v3 = v2 + 1
    if diagonal == "right":
        t0 = np.stack([v0, v1, v3], axis=1)
        t1 = np.stack([v0, v3, v2], axis=1)
    elif diagonal == "left":
        t0 = np.stack([v0, v1, v2], axis=1)
        t1 = np.stack([v1, v3, v2], axis=1)
    else:
        raise ValueError(diagonal)
    cells = np.empty((2 * nx * ny, 3), dtype=np.int32)
    cells[0::2] = t0
    cells[1::2] = t1
    return cells


def create_unit_square(comm, nx: int, ny: int, diagonal: str = "right", cell_type: str = "triangle") -> Mesh:
    """Same call shape as `dolfinx.mesh.create_unit_square(comm, nx, ny[, cell_type])`."""
    return create_rectangle((0.0, 0.0), (1.0, 1.0), nx, ny, diagonal, comm, cell_type)


def create_box(p0, p1, nx: int, ny: int, nz: int, comm=None) -> Mesh:
    """Structured tetrahedral mesh of a box: every grid cube is split into the six Kuhn tetrahedra
    around its main diagonal (conforming across cubes), vertices numbered x fastest.  Same role as
    `dolfinx.mesh.create_box(..., CellType.tetrahedron)`; DOLFINx's own numbering cannot be reproduced
    (SURVEY App. A), parity is defined relative to the arrays handed over."""
    xs = np.linspace(p0[0], p1[0], nx + 1)
    ys = np.linspace(p0[1], p1[1], ny + 1)
    zs = np.linspace(p0[2], p1[2], nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1)
    vid = lambda q: (q[:, 2] * (ny + 1) + q[:, 1]) * (nx + 1) + q[:, 0]
    cells = []
    for perm in ((0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)):
        v = [base.copy()]
        for ax in perm:
            w = v[-1].copy()
            w[:, ax] += 1
            v.append(w)
        cells.append(np.stack([vid(q) for q in v], axis=1))
    cells = np.stack(cells, axis=1).reshape(-1, 4).astype(np.int32)     # the six tetrahedra of a cube are consecutive
    return Mesh(pts, cells, comm)


def create_unit_cube(comm, nx: int, ny: int, nz: int) -> Mesh:
    """Same call shape as `dolfinx.mesh.create_unit_cube(comm, nx, ny, nz)` (tetrahedra; reference
    src/scenarios/taylor_green.py:34)."""
    return create_box((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), nx, ny, nz, comm)
