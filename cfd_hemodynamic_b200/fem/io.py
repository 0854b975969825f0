"""Mesh and solution files either side of the hot path.

`read_from_msh` stands in for `dolfinx.io.gmshio.read_from_msh(path, comm, rank, gdim)` as the
reference's experiment scenarios call it (src/experiments/scenario_factory.py:46-48: 3-D artery
meshes written by gmsh, physical groups INLET / OUTLET / WALL / FLUID): it returns
`(mesh, cell_tags, facet_tags)` with the 3P semantics restated from DOLFINx v0.9 —

* only elements that belong to a physical group are read; the cells of the mesh are the elements of
  the highest topological dimension present, `cell_tags` carries their physical tags;
* elements one dimension lower become `facet_tags` (matched to the mesh facets by their vertex sets);
* gmsh node order -> DOLFINx order: identity for lines, triangles and tetrahedra, (0, 1, 3, 2) for
  quadrilaterals (tensor order); nodes no cell refers to are dropped and the rest renumbered in file
  order (DOLFINx additionally reorders for locality — parity is defined relative to the arrays handed
  over).

Formats: MSH 4.1 and 2.2, ASCII (gmsh 4.15 writes 4.1 by default; `Mesh.Binary` is 0 by default and
the reference never sets it).  `write_msh` writes 4.1 or 2.2 ASCII (used by the tests and to hand our
generated meshes to a DOLFINx installation for cross-checks); `write_vtu` writes P1 / Q1 fields as
VTK XML unstructured grids for ParaView, the role of `VTXWriter` in src/scenario.py:208-215.
"""
from __future__ import annotations

import numpy as np

from .mesh import Mesh, MeshTags, meshtags

__all__ = ["read_from_msh", "write_msh", "write_vtu"]

# gmsh element type -> (topological dimension, nodes, DOLFINx cell name, permutation gmsh -> DOLFINx)
_GMSH = {
    15: (0, 1, "point", (0,)),
    1: (1, 2, "interval", (0, 1)),
    2: (2, 3, "triangle", (0, 1, 2)),
    3: (2, 4, "quadrilateral", (0, 1, 3, 2)),
    4: (3, 4, "tetrahedron", (0, 1, 2, 3)),
}
_TYPE_OF = {"interval": 1, "triangle": 2, "quadrilateral": 3, "tetrahedron": 4}


def _sections(path):
    out, name, buf = {}, None, []
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            if line.startswith("$End"):
                out[name] = buf
                name, buf = None, []
            elif line.startswith("$"):
                name, buf = line[1:], []
            elif name is not None:
                buf.append(line)
    return out


def _parse_v4(sec):
    """-> node tags (m,), coordinates (m, 3), list of (gmsh type, physical tag, node tags (k, nn))."""
    # entities: (dim, tag) -> first physical tag (gmshio uses one physical group per entity)
    phys = {}
    ent = sec.get("Entities", [])
    if ent:
        counts = [int(v) for v in ent[0].split()]
        row = 1
        for dim, cnt in enumerate(counts):
            for _ in range(cnt):
                t = ent[row].split()
                row += 1
                tag = int(t[0])
                off = 4 if dim == 0 else 7
                nphys = int(t[off])
                if nphys:
                    phys[(dim, tag)] = abs(int(t[off + 1]))
    nodes = sec["Nodes"]
    nblocks, nnodes = (int(v) for v in nodes[0].split()[:2])
    tags = np.empty(nnodes, dtype=np.int64)
    xyz = np.empty((nnodes, 3))
    row, k = 1, 0
    for _ in range(nblocks):
        _, _, parametric, nb = (int(v) for v in nodes[row].split())
        row += 1
        tags[k:k + nb] = [int(nodes[row + i]) for i in range(nb)]
        row += nb
        for i in range(nb):
            xyz[k + i] = [float(v) for v in nodes[row + i].split()[:3]]
        row += nb
        k += nb
    elems = sec["Elements"]
    nblocks = int(elems[0].split()[0])
    row = 1
    groups = []
    for _ in range(nblocks):
        edim, etag, etype, nb = (int(v) for v in elems[row].split())
        row += 1
        block = np.array([[int(v) for v in elems[row + i].split()[1:]] for i in range(nb)], dtype=np.int64)
        row += nb
        if (edim, etag) in phys and etype in _GMSH:
            groups.append((etype, phys[(edim, etag)], block))
    return tags, xyz, groups


def _parse_v2(sec):
    nodes = sec["Nodes"]
    nnodes = int(nodes[0])
    tags = np.empty(nnodes, dtype=np.int64)
    xyz = np.empty((nnodes, 3))
    for i in range(nnodes):
        t = nodes[1 + i].split()
        tags[i] = int(t[0])
        xyz[i] = [float(v) for v in t[1:4]]
    elems = sec["Elements"]
    by = {}
    for i in range(int(elems[0])):
        t = [int(v) for v in elems[1 + i].split()]
        etype, ntags = t[1], t[2]
        if etype not in _GMSH or ntags < 1 or t[3] == 0:
            continue                                     # no physical group
        by.setdefault((etype, t[3]), []).append(t[3 + ntags:])
    groups = [(etype, ptag, np.array(rows, dtype=np.int64)) for (etype, ptag), rows in by.items()]
    return tags, xyz, groups


def read_from_msh(path, comm=None, rank: int = 0, gdim: int = 3):
    """(mesh, cell_tags, facet_tags) from a gmsh file; see the module docstring for the semantics."""
    sec = _sections(str(path))
    version = float(sec["MeshFormat"][0].split()[0])
    if int(sec["MeshFormat"][0].split()[1]) != 0:
        raise NotImplementedError("binary .msh files are not supported (gmsh writes ASCII unless Mesh.Binary is set)")
    tags, xyz, groups = _parse_v4(sec) if version >= 4.0 else _parse_v2(sec)
    if not groups:
        raise ValueError(f"{path}: no element belongs to a physical group")
    tdim = max(_GMSH[etype][0] for etype, _, _ in groups)
    cell_types = {etype for etype, _, _ in groups if _GMSH[etype][0] == tdim}
    if len(cell_types) != 1:
        raise NotImplementedError("mixed-cell meshes are not supported")
    ctype = cell_types.pop()
    _, _, cname, perm = _GMSH[ctype]
    cells = np.concatenate([b for etype, _, b in groups if etype == ctype])[:, list(perm)]
    cvals = np.concatenate([np.full(len(b), p, dtype=np.int32) for etype, p, b in groups if etype == ctype])
    # node tags -> consecutive indices over the nodes the cells use, in file order
    index_of_tag = np.full(int(tags.max()) + 1, -1, dtype=np.int64)
    index_of_tag[tags] = np.arange(len(tags))
    used = np.zeros(len(tags), dtype=bool)
    used[index_of_tag[cells.reshape(-1)]] = True
    new_of_old = np.cumsum(used) - 1
    renum = lambda a: new_of_old[index_of_tag[a]]
    x = xyz[used]
    if tdim < 3 and np.abs(x[:, tdim:]).max() > 0.0:
        raise NotImplementedError("manifold meshes (gdim > tdim) are not supported")
    mesh = Mesh(x[:, :tdim], renum(cells).astype(np.int32), comm,
                cell_type="quadrilateral" if cname == "quadrilateral" else None)
    cell_tags = meshtags(mesh, tdim, np.arange(len(cvals), dtype=np.int32), cvals)
    # facets: elements of dimension tdim - 1, matched to mesh facets through their sorted vertex tuples
    fgroups = [(p, b) for etype, p, b in groups if _GMSH[etype][0] == tdim - 1]
    if fgroups:
        fv = np.concatenate([renum(b) for _, b in fgroups])
        fvals = np.concatenate([np.full(len(b), p, dtype=np.int32) for p, b in fgroups])
        nfv = fv.shape[1]
        # match by the sorted vertex tuples themselves (lexicographic rows): a packed integer key would
        # wrap beyond 2^63 for triangular facets of meshes with more than ~2.1 M vertices
        mv = np.sort(mesh.topology.facet_vertices, axis=1).astype(np.int64)
        tv = np.sort(fv, axis=1).astype(np.int64)
        both = np.concatenate([mv, tv])
        _, inv = np.unique(both, axis=0, return_inverse=True)
        inv = np.asarray(inv).reshape(-1)
        mesh_keys, tag_keys = inv[:len(mv)], inv[len(mv):]
        order = np.argsort(mesh_keys, kind="stable")
        pos = np.searchsorted(mesh_keys[order], tag_keys)
        if (pos >= len(order)).any() or (mesh_keys[order][np.minimum(pos, len(order) - 1)] != tag_keys).any():
            raise ValueError(f"{path}: a tagged facet is not a facet of any cell")
        facet_tags = meshtags(mesh, tdim - 1, order[pos].astype(np.int32), fvals)
    else:
        facet_tags = meshtags(mesh, tdim - 1, np.zeros(0, np.int32), np.zeros(0, np.int32))
    cell_tags.name, facet_tags.name = "Cell tags", "Facet tags"
    return mesh, cell_tags, facet_tags


def write_msh(path, mesh: Mesh, cell_tag: int = 1, facet_tags: MeshTags | None = None, version: str = "4.1"):
    """ASCII .msh with one physical group for the cells (`cell_tag`) and one per value of `facet_tags`.
    Nodes are written 1-based in mesh order; quadrilaterals go back to gmsh's counter-clockwise order."""
    cname = mesh.topology.cell_name()
    tdim = mesh.topology.dim
    ctype = _TYPE_OF[cname]
    inv = np.argsort(_GMSH[ctype][3])
    cells = mesh.geometry.dofmap[:, inv] + 1
    x = mesh.geometry.x
    fblocks = []
    if facet_tags is not None and len(facet_tags.indices):
        ftype = {1: 15, 2: 1, 3: 2}[tdim] if cname != "quadrilateral" else 1
        fv = mesh.topology.facet_vertices
        for val in np.unique(facet_tags.values):
            fblocks.append((int(val), ftype, fv[facet_tags.find(val)] + 1))
    with open(path, "w") as fh:
        if version.startswith("2"):
            fh.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % len(x))
            for i, p in enumerate(x):
                fh.write("%d %.17g %.17g %.17g\n" % (i + 1, p[0], p[1], p[2]))
            fh.write("$EndNodes\n$Elements\n%d\n" % (len(cells) + sum(len(b) for _, _, b in fblocks)))
            e = 1
            for val, ftype, b in fblocks:
                for row in b:
                    fh.write("%d %d 2 %d %d %s\n" % (e, ftype, val, val, " ".join(str(v) for v in row)))
                    e += 1
            for row in cells:
                fh.write("%d %d 2 %d %d %s\n" % (e, ctype, cell_tag, cell_tag, " ".join(str(v) for v in row)))
                e += 1
            fh.write("$EndElements\n")
            return
        # 4.1: one entity per physical group (entity tag = physical tag), bounding boxes of the whole mesh
        lo, hi = x.min(axis=0), x.max(axis=0)
        box = "%.17g %.17g %.17g %.17g %.17g %.17g" % (*lo, *hi)
        counts = [0, 0, 0, 0]
        counts[tdim] = 1
        counts[tdim - 1] = len(fblocks)
        fh.write("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n$Entities\n%d %d %d %d\n" % tuple(counts))
        for dim in range(4):
            if dim == tdim - 1:
                for val, _, _ in fblocks:
                    fh.write(("%d %.17g %.17g %.17g 1 %d\n" % (val, *lo, val)) if dim == 0 else "%d %s 1 %d 0\n" % (val, box, val))
            elif dim == tdim:
                fh.write("%d %s 1 %d 0\n" % (cell_tag, box, cell_tag))
        fh.write("$EndEntities\n$Nodes\n1 %d 1 %d\n%d %d 0 %d\n" % (len(x), len(x), tdim, cell_tag, len(x)))
        fh.write("\n".join(str(i + 1) for i in range(len(x))) + "\n")
        for p in x:
            fh.write("%.17g %.17g %.17g\n" % (p[0], p[1], p[2]))
        nel = len(cells) + sum(len(b) for _, _, b in fblocks)
        fh.write("$EndNodes\n$Elements\n%d %d 1 %d\n" % (1 + len(fblocks), nel, nel))
        e = 1
        for val, ftype, b in fblocks:
            fh.write("%d %d %d %d\n" % (tdim - 1, val, ftype, len(b)))
            for row in b:
                fh.write("%d %s\n" % (e, " ".join(str(v) for v in row)))
                e += 1
        fh.write("%d %d %d %d\n" % (tdim, cell_tag, ctype, len(cells)))
        for row in cells:
            fh.write("%d %s\n" % (e, " ".join(str(v) for v in row)))
            e += 1
        fh.write("$EndElements\n")


_VTK_TYPE = {"triangle": 5, "quadrilateral": 9, "tetrahedron": 10}


def write_vtu(path, mesh: Mesh, point_data: dict):
    """VTK XML unstructured grid (ASCII) with nodal fields: `point_data` maps a name to an array of
    length n (scalar) or bs*n (interleaved vector, padded to 3 components like ParaView expects)."""
    cname = mesh.topology.cell_name()
    cells = mesh.geometry.dofmap
    if cname == "quadrilateral":
        cells = cells[:, [0, 1, 3, 2]]                   # VTK_QUAD is counter-clockwise
    n, E, nv = mesh.num_vertices, cells.shape[0], cells.shape[1]
    fmt = lambda a: " ".join(repr(float(v)) for v in np.asarray(a).reshape(-1))
    with open(path, "w") as fh:
        fh.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian">\n')
        fh.write('<UnstructuredGrid>\n<Piece NumberOfPoints="%d" NumberOfCells="%d">\n<PointData>\n' % (n, E))
        for name, arr in point_data.items():
            arr = np.asarray(arr, dtype=np.float64)
            bs = arr.size // n
            if bs * n != arr.size:
                raise ValueError(f"field {name}: length {arr.size} is not a multiple of the number of nodes")
            if bs > 1:
                v = np.zeros((n, 3))
                v[:, :bs] = arr.reshape(n, bs)
                fh.write('<DataArray type="Float64" Name="%s" NumberOfComponents="3" format="ascii">\n%s\n</DataArray>\n' % (name, fmt(v)))
            else:
                fh.write('<DataArray type="Float64" Name="%s" format="ascii">\n%s\n</DataArray>\n' % (name, fmt(arr)))
        fh.write('</PointData>\n<Points>\n<DataArray type="Float64" NumberOfComponents="3" format="ascii">\n%s\n</DataArray>\n</Points>\n'
                 % fmt(mesh.geometry.x))
        fh.write('<Cells>\n<DataArray type="Int32" Name="connectivity" format="ascii">\n%s\n</DataArray>\n'
                 % " ".join(str(int(v)) for v in cells.reshape(-1)))
        fh.write('<DataArray type="Int32" Name="offsets" format="ascii">\n%s\n</DataArray>\n'
                 % " ".join(str(nv * (i + 1)) for i in range(E)))
        fh.write('<DataArray type="UInt8" Name="types" format="ascii">\n%s\n</DataArray>\n</Cells>\n'
                 % " ".join([str(_VTK_TYPE[cname])] * E))
        fh.write('</Piece>\n</UnstructuredGrid>\n</VTKFile>\n')
