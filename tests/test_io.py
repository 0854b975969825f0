"""Mesh / solution files either side of the hot path (fem/io.py): gmsh .msh reader with
the semantics of dolfinx.io.gmshio.read_from_msh (src/experiments/scenario_factory.py:46-48) on hand-written
MSH 4.1 / 2.2 files laid out the way gmsh writes them, round trips through write_msh, and the VTU writer."""
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from cfd_hemodynamic_b200.fem import io as IO
from cfd_hemodynamic_b200.fem import mesh as M

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_read_msh41_tetrahedra_with_physical_groups():
    mesh, ct, ft = IO.read_from_msh(os.path.join(GOLD, "two_tets_v41.msh"), None, 0, gdim=3)
    assert mesh.topology.cell_name() == "tetrahedron" and mesh.topology.dim == 3
    # node 9 is not referenced by any cell: dropped; tags 1, 2, 3, 5, 7 -> 0..4 in file order
    assert mesh.num_vertices == 5
    assert np.array_equal(mesh.geometry.x, [[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 1], [0, 0, 1]])
    assert np.array_equal(mesh.geometry.dofmap, [[0, 1, 2, 4], [1, 2, 4, 3]])
    assert ct.dim == 3 and np.array_equal(ct.values, [4, 4]) and np.array_equal(ct.indices, [0, 1])
    # surface 13 carries no physical group: its triangle is not tagged
    assert ft.dim == 2 and len(ft.indices) == 3
    fv = mesh.topology.facet_vertices
    tagged = {tuple(fv[i]): v for i, v in zip(ft.indices, ft.values)}
    assert tagged == {(0, 1, 2): 1, (1, 2, 3): 3, (0, 1, 4): 3}
    assert np.array_equal(ft.find(1), [i for i in ft.indices if tuple(fv[i]) == (0, 1, 2)])
    assert set(ft.indices) <= set(M.exterior_facet_indices(mesh.topology))


def test_read_msh22_quadrilaterals_are_tensor_ordered():
    mesh, ct, ft = IO.read_from_msh(os.path.join(GOLD, "two_quads_v22.msh"), None, 0, gdim=2)
    assert mesh.topology.cell_name() == "quadrilateral" and mesh.geometry.dim == 2
    # gmsh counter-clockwise (1, 2, 5, 4) -> DOLFINx tensor order (0, 1, 3, 4): vertices 0-1 bottom, 2-3 top
    assert np.array_equal(mesh.geometry.dofmap, [[0, 1, 3, 4], [1, 2, 4, 5]])
    X = mesh.geometry.x[mesh.geometry.dofmap][:, :, :2]
    assert np.allclose(X[:, 1] - X[:, 0], [1, 0]) and np.allclose(X[:, 2] - X[:, 0], [0, 1])
    assert np.array_equal(ct.values, [1, 1])
    # the line without a physical group (second tag list starts with 0) is skipped
    assert len(ft.indices) == 1 and ft.values[0] == 2
    assert sorted(mesh.topology.facet_vertices[ft.indices[0]]) == [0, 3]


@pytest.mark.parametrize("version", ["4.1", "2.2"])
@pytest.mark.parametrize("kind", ["triangle", "quadrilateral", "tetrahedron"])
def test_msh_round_trip(tmp_path, kind, version):
    if kind == "tetrahedron":
        mesh = M.create_unit_cube(None, 2, 3, 2)
        marker = lambda x: np.isclose(x[0], 0.0)
    else:
        mesh = M.create_unit_square(None, 3, 4, cell_type=kind)
        marker = lambda x: np.isclose(x[1], 1.0)
    fdim = mesh.topology.dim - 1
    ext = M.exterior_facet_indices(mesh.topology)
    sel = M.locate_entities_boundary(mesh, fdim, marker)
    vals = np.where(np.isin(ext, sel), 2, 3).astype(np.int32)
    ft = M.meshtags(mesh, fdim, ext, vals)
    path = str(tmp_path / f"{kind}.msh")
    IO.write_msh(path, mesh, cell_tag=7, facet_tags=ft, version=version)
    mesh2, ct2, ft2 = IO.read_from_msh(path, None, 0, gdim=mesh.geometry.dim)
    assert mesh2.topology.cell_name() == kind
    assert np.array_equal(mesh2.geometry.x, mesh.geometry.x)
    assert np.array_equal(mesh2.geometry.dofmap, mesh.geometry.dofmap)
    assert np.all(ct2.values == 7) and len(ct2.values) == mesh.num_cells
    # same facets, same tags (facet numbering is rebuilt from the cells, hence identical)
    assert np.array_equal(ft2.indices, ft.indices) and np.array_equal(ft2.values, ft.values)


def test_msh_errors(tmp_path):
    p = tmp_path / "bin.msh"
    p.write_text("$MeshFormat\n4.1 1 8\n$EndMeshFormat\n")
    with pytest.raises(NotImplementedError):
        IO.read_from_msh(str(p))
    p = tmp_path / "nogroups.msh"
    p.write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n3\n1 0 0 0\n2 1 0 0\n3 0 1 0\n$EndNodes\n"
                 "$Elements\n1\n1 2 2 0 1 1 2 3\n$EndElements\n")
    with pytest.raises(ValueError):
        IO.read_from_msh(str(p), gdim=2)


def test_write_vtu(tmp_path):
    mesh = M.create_unit_cube(None, 1, 1, 2)
    n = mesh.num_vertices
    u = np.arange(3 * n, dtype=float)
    p = np.linspace(0, 1, n)
    path = str(tmp_path / "out.vtu")
    IO.write_vtu(path, mesh, {"velocity": u, "pressure": p})
    root = ET.parse(path).getroot()
    piece = root.find("UnstructuredGrid/Piece")
    assert int(piece.get("NumberOfPoints")) == n and int(piece.get("NumberOfCells")) == 12
    arrays = {a.get("Name"): a for a in piece.iter("DataArray")}
    assert np.allclose(np.array(arrays["velocity"].text.split(), dtype=float), u)
    assert np.allclose(np.array(arrays["pressure"].text.split(), dtype=float), p)
    assert np.array_equal(np.array(arrays["connectivity"].text.split(), dtype=int), mesh.geometry.dofmap.reshape(-1))
    assert set(arrays["types"].text.split()) == {"10"}
    # 2-D vector fields are padded to three components; quadrilaterals leave in VTK's cyclic order
    quad = M.create_unit_square(None, 2, 1, cell_type="quadrilateral")
    IO.write_vtu(path, quad, {"velocity": np.ones(2 * quad.num_vertices)})
    arrays = {a.get("Name"): a for a in ET.parse(path).getroot().iter("DataArray")}
    v = np.array(arrays["velocity"].text.split(), dtype=float).reshape(-1, 3)
    assert np.all(v[:, :2] == 1.0) and np.all(v[:, 2] == 0.0)
    conn = np.array(arrays["connectivity"].text.split(), dtype=int).reshape(-1, 4)
    X = quad.geometry.x[conn][:, :, :2]
    area2 = sum(X[:, i, 0] * X[:, (i + 1) % 4, 1] - X[:, (i + 1) % 4, 0] * X[:, i, 1] for i in range(4))
    assert np.allclose(area2, 2 * 0.5)                       # counter-clockwise, no bow-tie
