"""oracle/pk_oracle.py (Pk-Pk triangles): k = 1 reproduces oracle/ns_oracle.py, k = 2 passes the Jacobian = dF/dx check by
complex step, the Poiseuille patch test (which needs the viscous part of the strong residual), the constant-pressure
null space, and a quadrature-independence check."""
import numpy as np

from cfd_hemodynamic_b200.fem import discretization as D
from cfd_hemodynamic_b200.fem import mesh as M
from cfd_hemodynamic_b200.fem import quadrature as Q
from oracle import ns_oracle as O
from oracle import pk_oracle as PK
from tests import common as T


def _p2_problem(mesh, dt=0.01, rho=1.3, mu=0.02, f=(0.3, -0.2), deg=8):
    x, cells6 = D.p2_nodes(mesh)
    rule = Q.triangle_rule(deg)
    rules = {k: rule for k in ("Fu", "Fp", "uu", "up", "pu", "pp")}
    return O.Problem(x=x, cells=cells6, h=PK.cell_diameter(x, cells6), dt=dt, rho=rho, mu=mu, f=np.asarray(f, float),
                     rules=rules, facet_rule=Q.interval_gauss(4))


def test_k1_reproduces_the_p1_oracle():
    mesh = T.perturbed_square(5, 4, seed=2)
    prob = T.make_problem(mesh)
    u, p, un = T.smooth_fields(prob.x)
    U, P, Un = O._gather(prob, u, p, un)
    rule = prob.rules["Fu"]
    for a, b in zip(PK.element_F(prob, U, P, Un, rule), O.element_F(prob, U, P, Un, rule)):
        assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max()
    for a, b in zip(PK.element_J(prob, U, P, Un, rule), O.element_J(prob, U, P, Un, rule)):
        assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max()
    ext = M.exterior_facet_indices(mesh.topology)
    fs = O.FacetSet(pairs=mesh.topology.facet_cell_pairs(ext), a_p=1.0, pconst=0.3, a_g=1.0, a_s=0.7, a_n=1.0, beta_n=50.0,
                    a_b=1.0, beta_b=0.2)
    ce = fs.pairs[:, 0]
    a, b = PK.facet_F(prob, fs, U[ce], P[ce], Un[ce]), O.facet_F(prob, fs, U[ce], P[ce], Un[ce])
    assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max()
    assert abs(PK.outlet_flux(prob, fs.pairs, un) - O.outlet_flux(prob, fs.pairs, un)) < 1e-13


def test_p2_jacobian_is_the_derivative_of_the_residual():
    mesh = T.perturbed_square(3, 3, seed=5)
    prob = _p2_problem(mesh)
    n = prob.n
    u, p, un = T.smooth_fields(prob.x)
    ext = M.exterior_facet_indices(mesh.topology)
    prob.facet_sets = [O.FacetSet(pairs=mesh.topology.facet_cell_pairs(ext), a_p=1.0, pconst=0.3, a_g=1.0, a_s=0.7, a_n=1.0,
                                  beta_n=50.0, a_b=1.0, beta_b=0.2)]
    A = O.assemble_J_raw(prob, u, p, un)
    rng = np.random.default_rng(0)
    d = rng.standard_normal(3 * n)
    hstep = 1e-30
    x = np.concatenate([u, p]).astype(complex) + 1j * hstep * d
    Fc = O.assemble_F_raw(prob, x[:2 * n], x[2 * n:], un)
    Jd = Fc.imag / hstep
    assert np.linalg.norm(A @ d - Jd) <= 1e-12 * np.linalg.norm(Jd)


def test_p2_poiseuille_patch_and_nullspace():
    """u = (4 y (1 - y), 0), p = -8 mu x solves the steady equations exactly and lies in P2 x P2: every interior row of
    the residual vanishes — SUPG / PSPG included, which requires the viscous part of the strong residual."""
    mesh = T.perturbed_square(4, 4, seed=7)
    mu = 0.05
    prob = _p2_problem(mesh, dt=1e9, rho=1.0, mu=mu, f=(0.0, 0.0))
    x = prob.x
    n = prob.n
    u = np.zeros((n, 2))
    u[:, 0] = 4.0 * x[:, 1] * (1.0 - x[:, 1])
    p = -8.0 * mu * x[:, 0]
    b = O.assemble_F_raw(prob, u.reshape(-1), p, u.reshape(-1))
    ext = M.exterior_facet_indices(mesh.topology)
    bnd_v = np.unique(mesh.topology.facet_vertices[ext])
    bnd = np.concatenate([bnd_v, mesh.geometry.x.shape[0] + ext])            # boundary vertex + edge nodes
    interior = np.setdiff1d(np.arange(n), bnd)
    bu, bp = b[:2 * n].reshape(-1, 2), b[2 * n:]
    scale = np.abs(bu).max()
    assert scale > 1e-3                                               # boundary rows carry the traction
    assert np.abs(bu[interior]).max() <= 1e-11 * scale
    assert np.abs(bp[interior]).max() <= 1e-11 * scale
    # constant pressure: in the kernel of J once the all-facet term of stabilized_schur.py:79 is present
    prob.facet_sets = [O.FacetSet(pairs=mesh.topology.facet_cell_pairs(ext), a_p=1.0, a_g=1.0)]
    A = O.assemble_J_raw(prob, u.reshape(-1), p, u.reshape(-1))
    c = np.zeros(3 * n)
    c[2 * n:] = 1.0
    assert np.abs(A @ c).max() <= 1e-11 * abs(A).max()


def test_p2_pattern_and_sizes():
    mesh = M.create_unit_square(None, 3, 2)
    prob = _p2_problem(mesh)
    nv, ne = mesh.geometry.x.shape[0], mesh.topology.facet_vertices.shape[0]
    assert prob.n == nv + ne and prob.cells.shape == (12, 6)
    rp, ci = O.sparsity_pattern(prob)
    assert rp.shape[0] == 3 * prob.n + 1 and (np.diff(ci.reshape(-1)) != 0).any()
