"""The curl-curl / rotational formulation (reference src/solvers/stabilized_schur_pressurebc.py:85-160,189-201) at the
level of the GLOBAL oracle: the element routines of oracle/curlcurl_oracle.py behind the kernel interfaces of
oracle/ns_oracle.py (2-D) and oracle/ns3d_oracle.py (3-D), so that assembly, Dirichlet treatment and Newton are shared.

* the assembled Jacobian is the derivative of the assembled residual (complex step through the whole assembly);
* known answer: plane Poiseuille flow driven by the weak pressure conditions of the form — with
  `pconst` on inlet / outlet the natural condition is the total pressure p + rho |u|^2 / 2 = pconst, the Nitsche terms
  hold u_T = 0 there, walls are no-slip — converges to u_max = dP H^2 / (8 mu L)."""
import numpy as np
import scipy.sparse.linalg as spla

from cfd_hemodynamic_b200.fem import mesh as M
from cfd_hemodynamic_b200.fem import quadrature as Q
from oracle import ns3d_oracle as O3
from oracle import ns_oracle as O
from oracle import simplex_oracle as S
from tests import common as T


def _channel(nx, ny, L=2.0, H=1.0):
    mesh = M.create_rectangle((0.0, 0.0), (L, H), nx, ny)
    x = mesh.geometry.x[:, :2]
    mid = lambda f: x[mesh.topology.facet_vertices[f]].mean(axis=1)
    ext = M.exterior_facet_indices(mesh.topology)
    fin = ext[np.isclose(mid(ext)[:, 0], 0.0)]
    fout = ext[np.isclose(mid(ext)[:, 0], L)]
    wall = np.nonzero(np.isclose(x[:, 1], 0.0) | np.isclose(x[:, 1], H))[0]
    return mesh, fin, fout, wall


def test_global_jacobian_is_the_derivative_of_the_residual_2d():
    mesh = T.perturbed_square(4, 3, seed=3)
    one = {k: Q.triangle_rule(8) for k in T.BLOCK_DEGREE}          # one rule for every block: J must equal dF/dx exactly
    prob = T.make_problem(mesh, rules=one)
    prob.formulation = "curlcurl"
    ext = M.exterior_facet_indices(mesh.topology)
    prob.facet_sets = [O.FacetSet(pairs=mesh.topology.facet_cell_pairs(ext[::2]), pconst=0.7, a_n=2.0, beta_n=30.0)]
    u, p, un = T.smooth_fields(prob.x)
    n = prob.n
    A = O.assemble_J_raw(prob, u, p, un).toarray()
    x0 = np.concatenate([u, p])
    step = 1e-30
    for j in range(0, 3 * n, 7):
        xc = x0.astype(complex)
        xc[j] += 1j * step
        col = O.assemble_F_raw(prob, xc[:2 * n], xc[2 * n:], un).imag / step
        assert np.abs(col - A[:, j]).max() <= 1e-12 * max(1.0, np.abs(A[:, j]).max())


def test_global_jacobian_is_the_derivative_of_the_residual_3d():
    from tests.test_tet_host import _perturbed_cube
    x, cells = _perturbed_cube(2, seed=5)
    n = x.shape[0]
    rule = S.tet_gauss_jacobi(6)
    rules = {k: rule for k in ("Fu", "Fp", "uu", "up", "pu", "pp")}
    pairs = S.exterior_facets(cells)
    prob = O3.Problem3D(x=x, cells=cells, dt=0.02, rho=1.1, mu=0.03, f=np.array([0.2, -0.1, 0.3]), rules=rules,
                        facet_sets=[O.FacetSet(pairs=pairs[::3], pconst=-0.4, a_n=1.0, beta_n=20.0)],
                        facet_rule=S.triangle_facet_rule(4), formulation="curlcurl")
    rng = np.random.default_rng(2)
    xk, un = rng.standard_normal(4 * n), rng.standard_normal(3 * n)
    A = O3.assemble_J_raw(prob, xk, un).toarray()
    step = 1e-30
    for j in range(0, 4 * n, 5):
        xc = xk.astype(complex)
        xc[j] += 1j * step
        col = O3.assemble_F_raw(prob, xc, un).imag / step
        assert np.abs(col - A[:, j]).max() <= 1e-12 * max(1.0, np.abs(A[:, j]).max())


def test_pressure_driven_poiseuille_known_answer():
    L, H, mu, rho, dP = 2.0, 1.0, 0.5, 1.0, 1.0
    mesh, fin, fout, wall = _channel(12, 10, L, H)
    prob = T.make_problem(mesh, dt=0.25, rho=rho, mu=mu, f=(0.0, 0.0))
    prob.formulation = "curlcurl"
    topo = mesh.topology
    prob.facet_sets = [O.FacetSet(pairs=topo.facet_cell_pairs(fin), pconst=dP, a_n=1.0, beta_n=100.0),
                       O.FacetSet(pairs=topo.facet_cell_pairs(fout), pconst=0.0, a_n=1.0, beta_n=100.0)]
    n = prob.n
    prob.bcs = T.oracle_bcs(prob, [("u", wall, np.zeros(2 * n))])
    x = np.zeros(3 * n)
    un = np.zeros(2 * n)
    for _ in range(30):                                   # implicit steps towards the steady state
        x_old = x
        x, its, reason = O.newton_solve(prob, x, un, rtol=1e-10, atol=1e-12)
        assert reason > 0
        un = x[:2 * n].copy()
    # the mid-point rule does not damp the 2 dt mode of the impulsive start: evaluate the mean of two consecutive steps
    assert np.abs(x - x_old)[:2 * n].max() < 0.05
    x = 0.5 * (x + x_old)
    u = x[:2 * n].reshape(-1, 2)
    umax = dP * H * H / (8.0 * mu * L)
    centre = np.isclose(prob.x[:, 1], 0.5 * H)
    mid = centre & (np.abs(prob.x[:, 0] - 0.5 * L) < 0.3 * L)
    assert np.abs(u[mid, 0] - umax).max() < 0.02 * umax           # P1, 10 cells across: within 2 % away from the open ends
    assert np.abs(u[centre, 0] - umax).max() < 0.12 * umax        # node-to-node wiggles of the weak conditions at the ends
    assert np.abs(u[:, 1]).max() < 0.03 * umax                    # parallel flow (2 % cross-flow in the end wiggles)
    p = x[2 * n:]
    xin = np.isclose(prob.x[:, 0], 0.0) & ~np.isclose(prob.x[:, 1], 0.0) & ~np.isclose(prob.x[:, 1], H)
    xout = np.isclose(prob.x[:, 0], L) & ~np.isclose(prob.x[:, 1], 0.0) & ~np.isclose(prob.x[:, 1], H)
    # static pressure drop = total pressure drop (same profile at both ends)
    assert abs((p[xin].mean() - p[xout].mean()) - dP) < 0.05 * dP
