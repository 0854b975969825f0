"""Duck-typed stand-ins for the live DOLFINx objects the reference hands to a solver plugin (test infrastructure for
integration/stabilized_schur_b200.py).  They reproduce what matters for a drop-in: the dof numbering differs from the
geometry node numbering (DOLFINx renumbers dofs), facets are numbered differently from the shim's own numbering, and
Dirichlet data comes as unrolled dof lists + value Functions living in the dof numbering."""
import numpy as np


class _X:
    def __init__(self, n):
        self.array = np.zeros(n)


class FakeFunction:
    def __init__(self, space):
        self.function_space = space
        self.x = _X(space.bs * space.n)
        self.name = "f"

    def interpolate(self, f):
        sp = self.function_space
        if isinstance(f, FakeFunction):
            self.x.array[:] = f.x.array
            return
        X = np.zeros((3, sp.n))
        X[:sp.coords.shape[1]] = sp.coords.T
        vals = np.asarray(f(X), dtype=float).reshape(sp.bs, sp.n)
        self.x.array[:] = vals.T.reshape(-1)


class _IndexMap:
    def __init__(self, n):
        self.size_local, self.num_ghosts = n, 0


class _Dofmap:
    def __init__(self, cells, bs, n):
        self.list = cells
        self.index_map_bs = bs
        self.index_map = _IndexMap(n)


class FakeSpace:
    def __init__(self, host, bs):
        self.mesh = host
        self.bs = bs
        self.n = host.n
        self.coords = host.x_dof
        self.dofmap = _Dofmap(host.cells_dof, bs, host.n)

    def tabulate_dof_coordinates(self):
        out = np.zeros((self.n, 3))
        out[:, :self.coords.shape[1]] = self.coords
        return out


class _Geometry:
    def __init__(self, x, cells, dim):
        self.x = np.hstack([x, np.zeros((x.shape[0], 3 - x.shape[1]))])
        self.dofmap = cells
        self.dim = dim


class _Topology:
    def __init__(self, shim_topology):
        self._t = shim_topology
        self.dim = shim_topology.dim

    def cell_name(self):
        return self._t.cell_name()

    def create_connectivity(self, *_):
        return None


class _Comm:
    size, rank = 1, 0


class FakeDirichletBC:
    def __init__(self, g, dofs_unrolled, refresh=None):
        self.g = g
        self._dofs = np.asarray(dofs_unrolled, dtype=np.int32)
        self._refresh = refresh

    def dof_indices(self):
        return self._dofs, len(self._dofs)

    def update(self):
        if self._refresh:
            self._refresh(self.g)


class FakeBoundaryCondition:
    """What src/boundaryCondition.py gives: getBC(V) -> DirichletBC with .update()."""

    def __init__(self, host, geom_nodes, value_fn, refresh=None):
        self.host, self.nodes, self.value_fn, self.refresh = host, np.asarray(geom_nodes), value_fn, refresh

    def getBC(self, V):
        g = FakeFunction(V)
        g.interpolate(self.value_fn)
        blocks = self.host.perm[self.nodes]
        dofs = (V.bs * blocks[:, None] + np.arange(V.bs)[None, :]).reshape(-1)
        return FakeDirichletBC(g, np.sort(dofs), self.refresh)


class FakeTags:
    def __init__(self, host, shim_tags):
        self.host = host
        self._t = shim_tags

    def find(self, value):
        return self.host.to_host_facets(self._t.find(value))


class FakeDolfinxHost:
    """A mesh + spaces as DOLFINx would present them, built from a shim mesh `m` and a dof permutation."""

    def __init__(self, m, seed=0):
        rng = np.random.default_rng(seed)
        self.shim = m
        gdim = m.geometry.dim
        x = m.geometry.x[:, :gdim]
        self.n = x.shape[0]
        self.perm = rng.permutation(self.n)                 # dof of geometry node i
        self.x_dof = np.empty_like(x)
        self.x_dof[self.perm] = x
        self.cells_dof = self.perm[m.geometry.dofmap].astype(np.int32)
        self.geometry = _Geometry(x, m.geometry.dofmap, gdim)
        self.topology = _Topology(m.topology)
        self.comm = _Comm()
        self.nf = m.topology.facet_vertices.shape[0]
        self.V = FakeSpace(self, gdim)
        self.Q = FakeSpace(self, 1)

    def h(self, dim, cells):
        return self.shim.h(dim, cells)

    # facets are numbered backwards on the "DOLFINx" side
    def to_host_facets(self, shim_facets):
        return (self.nf - 1 - np.asarray(shim_facets, dtype=np.int64))[::-1].copy()

    def integration_entities(self, _mesh, host_facets):
        shim_f = self.nf - 1 - np.asarray(host_facets, dtype=np.int64)
        return self.shim.topology.facet_cell_pairs(shim_f)

    def host_objects(self):
        V, Q = self.V, self.Q
        return dict(_V=V, _Q=Q, V=V, Q=Q, u_sol=FakeFunction(V), p_sol=FakeFunction(Q), u_prev=FakeFunction(V),
                    p_prev=FakeFunction(Q), u_residual=FakeFunction(V), p_residual=FakeFunction(Q))
