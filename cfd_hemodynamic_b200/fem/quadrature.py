"""Quadrature rules, passed to the kernels at run time (never compiled in).

The reference gets its rules from Basix through FFCx at `form(...)`
(reference src/solvers/stabilized_schur.py:188-189): the estimated degree of
each block form selects the default Basix scheme — Xiao–Gimbutas on simplices
up to degree 30, Gauss–Jacobi otherwise.  Basix is not in
this image, so the XG tables themselves are unavailable.  What ships here:

* `triangle_symmetric(degree)`: fully symmetric (D3-orbit) rules with the
  same point counts as the XG rules of the degrees the hot path needs
  (10 → 25, 11 → 28, 12 → 33 points), stored in `_tables.py` as orbit
  parameters (degrees 10, 11: found by `tools/make_triangle_rules.py` from the
  moment equations; degree 12: Dunavant's rule polished the same way) and
  expanded/verified here.
* `triangle_gauss_jacobi(degree)`: collapsed Gauss–Jacobi tensor rule (what
  Basix uses above degree 30) for any degree — the fallback.
* `interval_gauss(npts)`: Gauss–Legendre on [0, 1] for exterior facets.

Because the SUPG parameter is not polynomial, swapping a rule for another of
the same degree moves matrix entries by ~1e-6..1e-9 relative; the rule is an
explicit input of both the oracle and the CUDA kernels so the true Basix
tables can be injected when a DOLFINx install is at hand.
"""
from __future__ import annotations

import itertools
import math

import numpy as np

__all__ = [
    "triangle_symmetric", "triangle_gauss_jacobi", "interval_gauss",
    "triangle_rule", "check_triangle_rule", "expand_orbits", "quadrilateral_rule",
    "FACET_POINTS_QUAD",
]

# Facet rule on quadrilateral meshes: the traces of Q1 functions are linear, the backflow term
# (u_n.n)_- (u_m.v) is cubic apart from the kink and grad(u) is rational on non-affine cells;
# 4 Gauss points (degree 7) for every tagged ds integral.
FACET_POINTS_QUAD = 4


def interval_gauss(npts: int):
    """Gauss–Legendre on [0,1]; npts = (degree + 2)//2 (3P Basix GJ rule)."""
    x, w = np.polynomial.legendre.leggauss(npts)
    return 0.5 * (x + 1.0), 0.5 * w


def quadrilateral_rule(degree: int):
    """Default Basix scheme on the reference quadrilateral [0,1]^2: tensor product of the
    m-point Gauss–Jacobi (alpha = 0, i.e. Gauss–Legendre) rule, m = (degree + 2) // 2;
    first coordinate slowest; weights sum to 1."""
    m = (degree + 2) // 2
    x, w = interval_gauss(m)
    pts = np.stack([np.repeat(x, m), np.tile(x, m)], axis=1)
    wts = np.repeat(w, m) * np.tile(w, m)
    return pts, wts


def _gauss_jacobi(n: int, alpha: float):
    """Gauss–Jacobi nodes/weights on [0,1] for weight (1-x)^alpha."""
    from scipy.special import roots_jacobi
    x, w = roots_jacobi(n, alpha, 0.0)
    return 0.5 * (x + 1.0), w / 2.0 ** (alpha + 1.0)


def triangle_gauss_jacobi(degree: int):
    """Collapsed-coordinate rule on the reference triangle {x,y>=0, x+y<=1};
    weights sum to 1/2."""
    m = (degree + 2) // 2
    px, wx = _gauss_jacobi(m, 1.0)
    py, wy = _gauss_jacobi(m, 0.0)
    pts = np.empty((m * m, 2))
    wts = np.empty(m * m)
    k = 0
    for i in range(m):
        for j in range(m):
            pts[k, 0] = px[i]
            pts[k, 1] = py[j] * (1.0 - px[i])
            wts[k] = wx[i] * wy[j]
            k += 1
    return pts, wts


def tetrahedron_rule(degree: int):
    """Collapsed-coordinate Gauss–Jacobi rule on the reference tetrahedron {x,y,z>=0, x+y+z<=1}
    (weights sum to 1/6, ((degree + 2) // 2)^3 points, exact to `degree`).  Basix picks a
    Xiao–Gimbutas scheme with fewer points for these degrees (122 points at degree 12); its tables are
    not available here, and the rule is a run-time input of the library (`quadrature=` on the solver)."""
    m = (degree + 2) // 2
    px, wx = _gauss_jacobi(m, 2.0)
    py, wy = _gauss_jacobi(m, 1.0)
    pz, wz = _gauss_jacobi(m, 0.0)
    X, Y, Z = np.meshgrid(px, py, pz, indexing="ij")
    WX, WY, WZ = np.meshgrid(wx, wy, wz, indexing="ij")
    pts = np.stack([X, Y * (1.0 - X), Z * (1.0 - X) * (1.0 - Y)], axis=-1).reshape(-1, 3)
    return np.ascontiguousarray(pts), np.ascontiguousarray((WX * WY * WZ).reshape(-1))


def expand_orbits(centroid_w, s21, s111):
    """Expand D3 orbits to points (x, y) on the reference triangle and weights
    (sum 1/2).  s21: [(w, a)] → barycentric permutations of (a, a, 1-2a);
    s111: [(w, a, b)] → permutations of (a, b, 1-a-b)."""
    pts, wts = [], []
    if centroid_w is not None:
        pts.append((1.0 / 3.0, 1.0 / 3.0))
        wts.append(centroid_w)
    for w, a in s21:
        c = 1.0 - 2.0 * a
        for bary in ((a, a, c), (a, c, a), (c, a, a)):
            pts.append((bary[1], bary[2]))
            wts.append(w)
    for w, a, b in s111:
        c = 1.0 - a - b
        for bary in itertools.permutations((a, b, c)):
            pts.append((bary[1], bary[2]))
            wts.append(w)
    return np.array(pts, dtype=np.float64), np.array(wts, dtype=np.float64)


def check_triangle_rule(pts, wts, degree: int) -> float:
    """Max abs error of the rule on all monomials x^i y^j, i+j <= degree.
    Exact: i! j! / (i+j+2)!."""
    err = 0.0
    for i in range(degree + 1):
        for j in range(degree + 1 - i):
            exact = math.factorial(i) * math.factorial(j) / math.factorial(i + j + 2)
            q = float(np.sum(wts * pts[:, 0] ** i * pts[:, 1] ** j))
            err = max(err, abs(q - exact))
    return err


def triangle_symmetric(degree: int):
    from ._tables import TRIANGLE_ORBITS
    if degree not in TRIANGLE_ORBITS:
        raise KeyError(degree)
    t = TRIANGLE_ORBITS[degree]
    return expand_orbits(t["centroid"], t["s21"], t["s111"])


def triangle_rule(degree: int):
    """Default cell rule for a block form of estimated degree `degree`."""
    try:
        return triangle_symmetric(degree)
    except (KeyError, ImportError):
        return triangle_gauss_jacobi(degree)
