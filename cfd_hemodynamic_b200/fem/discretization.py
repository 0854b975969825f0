"""Host-side, one-time tables the CUDA library needs: node graph, exterior-facet
sets grouped by cell, Dirichlet flags.

3P behaviour restated: `create_matrix_block` builds the
full FE sparsity pattern (trigger: reference src/solvers/stabilized_schur.py:191);
for P1–P1 that pattern is the node adjacency graph expanded to the
[u interleaved | p] layout, which the library does on the device
(`hemo_get_pattern`).  Here we only build the scalar node graph.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .mesh import Mesh, exterior_facet_indices


def node_graph(cells: np.ndarray, n_nodes: int):
    """Sorted CSR adjacency (diagonal included) of the P1 dof graph."""
    c = cells.astype(np.int32)
    nv = c.shape[1]
    rows = np.repeat(c, nv, axis=1).reshape(-1)
    cols = np.tile(c, (1, nv)).reshape(-1)
    A = sp.coo_matrix((np.ones(rows.shape[0], dtype=np.int8), (rows, cols)), shape=(n_nodes, n_nodes)).tocsr()
    A.sort_indices()
    return A.indptr.astype(np.int32), A.indices.astype(np.int32)


def facet_set_by_cell(mesh: Mesh, facets: np.ndarray):
    """Group exterior facets by their cell: (cells (m,), mask (m,)) with bit
    `lf` of mask set when local facet lf of the cell is in the set."""
    pairs = mesh.topology.facet_cell_pairs(np.asarray(facets))
    if pairs.shape[0] == 0:
        return np.zeros(0, np.int32), np.zeros(0, np.int32)
    order = np.argsort(pairs[:, 0], kind="stable")
    pc = pairs[order, 0]
    bits = (1 << pairs[order, 1]).astype(np.int32)
    cells, start = np.unique(pc, return_index=True)
    mask = np.add.reduceat(bits, start).astype(np.int32)
    return cells.astype(np.int32), mask


def all_exterior_facets(mesh: Mesh) -> np.ndarray:
    return exterior_facet_indices(mesh.topology)


def pairs_by_cell(pairs: np.ndarray):
    """Group (cell, local facet) pairs by cell: (cells (m,), mask (m,)), bit lf of mask = facet lf."""
    pairs = np.asarray(pairs).reshape(-1, 2)
    if pairs.shape[0] == 0:
        return np.zeros(0, np.int32), np.zeros(0, np.int32)
    order = np.argsort(pairs[:, 0], kind="stable")
    pc = pairs[order, 0]
    bits = (1 << pairs[order, 1]).astype(np.int32)
    cells, start = np.unique(pc, return_index=True)
    return cells.astype(np.int32), np.add.reduceat(bits, start).astype(np.int32)


def dirichlet_arrays(n_nodes: int, cells: np.ndarray, bcs, gdim: int = 2):
    """bcs: list of (block 'u'|'p', block dof indices, value array (block-sized)).
    Returns dofflag (N uint8), dofmult (N f64), cellflag (E uint8), g (N f64) with
    N = (gdim + 1) n in the [u interleaved (gdim n) | p (n)] layout.
    Semantics (3P): diagonal += 1 per DirichletBC containing the
    dof; on shared dofs the last BC in the list provides the value."""
    N = (gdim + 1) * n_nodes
    nu = gdim * n_nodes
    flag = np.zeros(N, dtype=np.uint8)
    mult = np.zeros(N, dtype=np.float64)
    g = np.zeros(N, dtype=np.float64)
    for block, nodes, values in bcs:
        nodes = np.asarray(nodes, dtype=np.int64)
        if block == "u":
            d = (gdim * nodes[:, None] + np.arange(gdim)[None, :]).reshape(-1)
            vals = np.asarray(values, dtype=np.float64)[d]
        else:
            d = nu + nodes
            vals = np.asarray(values, dtype=np.float64)[nodes]
        flag[d] = 1
        mult[d] += 1.0
        g[d] = vals
    nodeflag = flag[nu:].astype(bool)
    for k in range(gdim):
        nodeflag |= flag[k:nu:gdim].astype(bool)
    cellflag = nodeflag[cells].any(axis=1).astype(np.uint8)
    return flag, mult, cellflag, g


def p2_nodes(mesh: Mesh):
    """Node set and cell table of the P2 Lagrange space on a triangle mesh (3P: Basix / DOLFINx dof layout):
    one node per vertex, then one per edge (= facet in 2-D) at its mid-point; a cell lists its three vertex nodes and
    then the nodes of the edges opposite local vertices 0, 1, 2.  Returns (x (Nv + Ne, 2), cells6 (E, 6) int32)."""
    if mesh.topology.cell_name() != "triangle":
        raise NotImplementedError("P2 node layout is implemented for triangles")
    x = mesh.geometry.x[:, :2]
    nv = x.shape[0]
    fv = mesh.topology.facet_vertices
    xe = 0.5 * (x[fv[:, 0]] + x[fv[:, 1]])
    c2f = mesh.topology.cell_facets
    cells6 = np.hstack([mesh.geometry.dofmap, nv + c2f]).astype(np.int32)
    return np.ascontiguousarray(np.vstack([x, xe])), np.ascontiguousarray(cells6)
