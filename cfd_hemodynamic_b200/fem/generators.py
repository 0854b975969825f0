"""Synthetic mesh generators for the BASELINE configs (no gmsh in this image).

Geometry and tags follow the reference scenarios:
  * `stenosis_structured`  — src/scenarios/stenosis_pressure_structured.py:190-393
    (transfinite grid between two walls made of line + 2 cubic Béziers + line;
    the reference recombines to quadrilaterals: `cell_type="quadrilateral"`;
    `"triangle"` splits every quad into two triangles, mirrored about the centre
    line — SURVEY.md §7.3-2)
  * `dfg_cylinder`         — src/scenarios/dfg_1.py:97-171 (channel 2.2 x 0.41,
    cylinder (0.2, 0.2) r = 0.05, graded size field LcMin near the cylinder)
Facet markers: inlet 2, outlet 3, wall 4, obstacle 5 (dfg_1.py:18-22).
"""
from __future__ import annotations

import math

import numpy as np

from .mesh import Mesh, MeshTags, exterior_facet_indices

INLET, OUTLET, WALL, OBSTACLE = 2, 3, 4, 5

STENOSIS_GRADES = {
    "mild": {"severity": 0.25, "slope": 0.3},
    "moderate": {"severity": 0.50, "slope": 0.3},
    "severe": {"severity": 0.75, "slope": 0.3},
}


def _bezier(p0, p1, p2, p3, n):
    """n+1 points, uniform in arc length, on the cubic Bézier p0..p3."""
    t = np.linspace(0.0, 1.0, 20 * n + 1)[:, None]
    c = ((1 - t) ** 3) * p0 + 3 * ((1 - t) ** 2) * t * p1 + 3 * (1 - t) * t ** 2 * p2 + t ** 3 * p3
    s = np.concatenate([[0.0], np.cumsum(np.linalg.norm(np.diff(c, axis=0), axis=1))])
    target = np.linspace(0.0, s[-1], n + 1)
    return np.stack([np.interp(target, s, c[:, 0]), np.interp(target, s, c[:, 1])], axis=1)


def stenosis_wall_points(L=138.0, R_in=1.57, R_out=1.2, res=0.15, x_position_stenosis=30.0, severity=0.75,
                         slope=0.3, tension=0.5, nx_inlet=None, nx_bezier=None, nx_line=None):
    """Top and bottom wall point chains of the transfinite grid and the inlet count
    (stenosis_pressure_structured.py:202-283)."""
    x_sten = x_position_stenosis
    r_taper_mid = R_in + (R_out - R_in) * (x_sten / L)
    R_min = (1.0 - severity) * r_taper_mid
    if R_min <= 0:
        raise ValueError("severity too large: stenosis would close the channel")
    h_sten = r_taper_mid - R_min
    dist_x = h_sten / slope if slope > 0 else L / 4
    dist_x = min(dist_x, min(x_sten, L - x_sten) * 0.95)
    cp1_x, cp2_x = x_sten - dist_x, x_sten + dist_x
    cp1_r = R_in + (R_out - R_in) * (cp1_x / L)
    cp2_r = R_in + (R_out - R_in) * (cp2_x / L)
    slope_top = (R_out - R_in) / L
    ha = hb = tension * dist_x
    if nx_inlet is not None:
        n_inlet = int(nx_inlet)
    else:
        n_inlet = max(4, int(math.ceil(2.0 * R_in / res)))
    if n_inlet % 2:
        n_inlet += 1
    n_bezier = max(2, int(nx_bezier)) if nx_bezier is not None else max(2, int(math.ceil(dist_x / res)))
    y_top_cp1, y_top_mid, y_top_cp2 = R_in + cp1_r, R_in + R_min, R_in + cp2_r
    len_pre = math.hypot(cp1_x, y_top_cp1 - 2.0 * R_in)
    len_post = math.hypot(L - cp2_x, (R_in + R_out) - y_top_cp2)
    if nx_line is not None:
        n_pre = n_post = max(2, int(nx_line))
    else:
        n_pre = max(2, int(math.ceil(len_pre / res)))
        n_post = max(2, int(math.ceil(len_post / res)))
    # top wall: line, Bézier, Bézier, line — radius r(x) about the centre line y = R_in
    A = lambda x, y: np.array([x, y])
    seg0 = np.linspace(A(0.0, 2.0 * R_in), A(cp1_x, y_top_cp1), n_pre + 1)
    seg1 = _bezier(A(cp1_x, y_top_cp1), A(cp1_x + ha, y_top_cp1 + ha * slope_top),
                   A(x_sten - hb, y_top_mid - hb * slope_top), A(x_sten, y_top_mid), n_bezier)
    seg2 = _bezier(A(x_sten, y_top_mid), A(x_sten + hb, y_top_mid + hb * slope_top),
                   A(cp2_x - ha, y_top_cp2 - ha * slope_top), A(cp2_x, y_top_cp2), n_bezier)
    seg3 = np.linspace(A(cp2_x, y_top_cp2), A(L, R_in + R_out), n_post + 1)
    top = np.concatenate([seg0[:-1], seg1[:-1], seg2[:-1], seg3])
    bot = top.copy()
    bot[:, 1] = 2.0 * R_in - top[:, 1]          # mirror about y = R_in
    return top, bot, n_inlet


def stenosis_structured(grade="severe", comm=None, cell_type="triangle", **mesh_options):
    """Transfinite stenosis channel + facet tags.  `cell_type="quadrilateral"`: the recombined
    Q1 mesh of the reference (stenosis_pressure_structured.py:379-386, tensor-ordered cells);
    `"triangle"`: every quad split into two P1 triangles."""
    opts = dict(L=138.0, R_in=1.57, R_out=1.2, res=0.15, x_position_stenosis=30.0, severity=0.567, slope=0.4,
                tension=0.5)
    opts.update(STENOSIS_GRADES.get(grade, STENOSIS_GRADES["severe"]))
    opts.update(mesh_options)
    top, bot, ny = stenosis_wall_points(**opts)
    nx = top.shape[0] - 1
    v = np.linspace(0.0, 1.0, ny + 1)
    # straight inlet/outlet sides: the transfinite (Coons) map reduces to a ruled surface
    X = (1.0 - v)[:, None, None] * bot[None, :, :] + v[:, None, None] * top[None, :, :]   # (ny+1, nx+1, 2)
    pts = X.reshape(-1, 2)
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v0 = (iy * (nx + 1) + ix).ravel()
    v1, v2 = v0 + 1, v0 + (nx + 1)
    v3 = v2 + 1
    if cell_type == "quadrilateral":
        mesh = Mesh(pts, np.stack([v0, v1, v2, v3], axis=1).astype(np.int32), comm, cell_type)
    elif cell_type == "triangle":
        lower = (iy.ravel() < ny // 2)
        cells = np.empty((2 * nx * ny, 3), dtype=np.int32)
        # mirrored diagonals: "right" below the centre line, "left" above
        cells[0::2] = np.where(lower[:, None], np.stack([v0, v1, v3], 1), np.stack([v0, v1, v2], 1))
        cells[1::2] = np.where(lower[:, None], np.stack([v0, v3, v2], 1), np.stack([v1, v3, v2], 1))
        mesh = Mesh(pts, cells, comm)
    else:
        raise ValueError(f"cell_type {cell_type!r}")
    L = opts["L"]
    ext = exterior_facet_indices(mesh.topology)
    fv = mesh.topology.facet_vertices[ext]
    xm = 0.5 * (pts[fv[:, 0], 0] + pts[fv[:, 1], 0])
    tol = 1e-9 * L
    vals = np.where(xm < tol, INLET, np.where(xm > L - tol, OUTLET, WALL)).astype(np.int32)
    ft = MeshTags(mesh, 1, ext, vals)
    mesh.mesh_options = opts
    return mesh, ft


def dfg_cylinder(lc_min=None, lc_max=None, comm=None, smooth_iters=8):
    """Graded triangle mesh of the DFG channel with cylinder by Delaunay
    triangulation of a graded point cloud (rings around the cylinder, staggered
    columns elsewhere) that follows the reference size field, dfg_1.py:146-155
    (Threshold: LcMin = r/6 within DistMin = r, LcMax = H/13 at DistMax = 2H)."""
    from scipy.spatial import Delaunay
    L, H, cx, cy, r = 2.2, 0.41, 0.2, 0.2, 0.05
    lc_min = r / 6.0 if lc_min is None else lc_min
    lc_max = H / 13.0 if lc_max is None else lc_max

    def lc_of(dist):
        """gmsh Threshold field: LcMin for dist <= DistMin = r, linear up to LcMax at DistMax = 2H."""
        t = np.clip((dist - r) / (2 * H - r), 0.0, 1.0)
        return lc_min + (lc_max - lc_min) * t

    pts = []
    fixed = []
    # cylinder boundary + rings following the size field
    radius, i = r, 0
    R1 = 0.145
    while radius < R1:
        l = float(lc_of(radius - r))
        m = max(12, int(round(2 * math.pi * radius / l)))
        th = 2 * math.pi * (np.arange(m) + 0.5 * (i % 2)) / m
        pts.append(np.stack([cx + radius * np.cos(th), cy + radius * np.sin(th)], axis=1))
        fixed.append(np.full(m, i == 0))
        radius += 0.866 * l
        i += 1
    r_out = radius - 0.4 * float(lc_of(radius - r))
    # outer boundary and interior columns whose spacing follows the size field in x
    xs = [0.0]
    while xs[-1] < L:
        l = float(lc_of(abs(xs[-1] - cx) - r))
        xs.append(xs[-1] + 0.866 * l)
    xs = np.array(xs) * (L / xs[-1])
    for k, xc in enumerate(xs):
        l = float(lc_of(abs(xc - cx) - r))
        ny = max(2, int(round(H / l)))
        if k % 2 and 0 < k < len(xs) - 1:
            ys = np.concatenate([[0.0], (np.arange(ny) + 0.5) * (H / ny), [H]])
        else:
            ys = np.linspace(0.0, H, ny + 1)
        col = np.stack([np.full_like(ys, xc), ys], axis=1)
        on_bnd = (ys == 0.0) | (ys == H) | (k == 0) | (k == len(xs) - 1)
        far = np.hypot(col[:, 0] - cx, col[:, 1] - cy) > r_out
        pts.append(col[far])
        fixed.append(on_bnd[far] if np.ndim(on_bnd) else np.full(int(far.sum()), bool(on_bnd)))
    P = np.concatenate(pts)
    fixed = np.concatenate(fixed)
    # drop ring points outside the channel or too close to its walls
    inside = (P[:, 0] > -1e-12) & (P[:, 0] < L + 1e-12) & (P[:, 1] > -1e-12) & (P[:, 1] < H + 1e-12)
    near_wall = (~fixed) & ((P[:, 1] < 0.5 * lc_min) | (P[:, 1] > H - 0.5 * lc_min))
    keep = inside & ~near_wall
    P, fixed = P[keep], fixed[keep]

    def triangulate(P):
        tri = Delaunay(P).simplices
        c = P[tri].mean(axis=1)
        ok = np.hypot(c[:, 0] - cx, c[:, 1] - cy) > r * (1 - 1e-9)
        tri = tri[ok]
        X = P[tri]
        area = 0.5 * ((X[:, 1, 0] - X[:, 0, 0]) * (X[:, 2, 1] - X[:, 0, 1]) - (X[:, 2, 0] - X[:, 0, 0]) * (X[:, 1, 1] - X[:, 0, 1]))
        return tri[np.abs(area) > 1e-14]

    tri = triangulate(P)
    for _ in range(smooth_iters):          # Laplacian smoothing of free nodes, then re-triangulate
        n = P.shape[0]
        acc = np.zeros_like(P)
        cnt = np.zeros(n)
        for a, b in ((0, 1), (1, 2), (2, 0)):
            np.add.at(acc, tri[:, a], P[tri[:, b]])
            np.add.at(acc, tri[:, b], P[tri[:, a]])
            np.add.at(cnt, tri[:, a], 1)
            np.add.at(cnt, tri[:, b], 1)
        newP = np.where(fixed[:, None] | (cnt[:, None] == 0), P, acc / np.maximum(cnt, 1)[:, None])
        P = 0.5 * P + 0.5 * newP
        tri = triangulate(P)
    used = np.unique(tri)
    remap = -np.ones(P.shape[0], dtype=np.int64)
    remap[used] = np.arange(used.shape[0])
    P = P[used]
    tri = remap[tri].astype(np.int32)
    mesh = Mesh(P, tri, comm)
    ext = exterior_facet_indices(mesh.topology)
    fv = mesh.topology.facet_vertices[ext]
    mid = 0.5 * (P[fv[:, 0]] + P[fv[:, 1]])
    vals = np.full(ext.shape[0], WALL, dtype=np.int32)
    vals[mid[:, 0] < 1e-9] = INLET
    vals[mid[:, 0] > L - 1e-9] = OUTLET
    vals[np.hypot(mid[:, 0] - cx, mid[:, 1] - cy) < 1.5 * r] = OBSTACLE
    ft = MeshTags(mesh, 1, ext, vals)
    return mesh, ft
