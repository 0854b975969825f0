"""3-D known-answer test of the tetrahedral oracle: the Ethier–Steinman solution of the
incompressible Navier–Stokes equations (nu = 1, f = 0) that the reference's taylor_green scenario
compares with (src/scenarios/taylor_green.py:74-134), Dirichlet velocity and pressure on the whole
boundary like there (:41-58).  The nodal velocity error must fall with mesh refinement."""
import numpy as np

from oracle import ns3d_oracle as N3
from oracle import simplex_oracle as S

A_, D_ = np.pi / 4, np.pi / 2


def exact_velocity(x, t):
    X, Y, Z = x[:, 0], x[:, 1], x[:, 2]
    e = np.exp(-D_ * D_ * t)
    return np.stack([
        -A_ * (np.exp(A_ * X) * np.sin(A_ * Y + D_ * Z) + np.exp(A_ * Z) * np.cos(A_ * X + D_ * Y)) * e,
        -A_ * (np.exp(A_ * Y) * np.sin(A_ * Z + D_ * X) + np.exp(A_ * X) * np.cos(A_ * Y + D_ * Z)) * e,
        -A_ * (np.exp(A_ * Z) * np.sin(A_ * X + D_ * Y) + np.exp(A_ * Y) * np.cos(A_ * Z + D_ * X)) * e], axis=1)


def exact_pressure(x, t):
    X, Y, Z = x[:, 0], x[:, 1], x[:, 2]
    return (-A_ * A_ / 2 * (np.exp(2 * A_ * X) + np.exp(2 * A_ * Y) + np.exp(2 * A_ * Z)
                            + 2 * np.sin(A_ * X + D_ * Y) * np.cos(A_ * Z + D_ * X) * np.exp(A_ * (Y + Z))
                            + 2 * np.sin(A_ * Y + D_ * Z) * np.cos(A_ * X + D_ * Y) * np.exp(A_ * (Z + X))
                            + 2 * np.sin(A_ * Z + D_ * X) * np.cos(A_ * Y + D_ * Z) * np.exp(A_ * (X + Y)))
            * np.exp(-2 * D_ * D_ * t))


def _run(n, steps=4, dt=0.005):
    x, cells = N3.unit_cube_tets(n)
    rules = {k: S.tet_gauss_jacobi(deg) for k, deg in dict(Fu=6, Fp=5, uu=6, up=5, pu=5, pp=4).items()}
    prob = N3.Problem3D(x=x, cells=cells, dt=dt, rho=1.0, mu=1.0, f=np.zeros(3), rules=rules)
    nn = prob.n
    boundary = np.nonzero((np.abs(x - 0.5) > 0.5 - 1e-12).any(axis=1))[0]
    prob.bc_dofs = np.concatenate([(3 * boundary[:, None] + np.arange(3)[None, :]).reshape(-1), 3 * nn + boundary])
    un = exact_velocity(x, 0.0).reshape(-1)
    xk = np.concatenate([un, exact_pressure(x, 0.0)])
    t = 0.0
    for _ in range(steps):
        t += dt
        g = np.concatenate([exact_velocity(x, t).reshape(-1), exact_pressure(x, t)])
        xk, its = N3.newton_step(prob, xk, un, g)
        assert its < 8
        un = xk[:3 * nn].copy()
    ue = exact_velocity(x, t).reshape(-1)
    return np.linalg.norm(un - ue) / np.linalg.norm(ue), prob


def test_ethier_steinman_error_decreases_with_refinement():
    e4, _ = _run(4)
    e8, prob = _run(8)
    assert e8 < 2e-3, e8
    assert e8 < 0.45 * e4, (e4, e8)                     # between first and second order on these coarse meshes


def test_3d_jacobian_is_derivative_of_residual_with_dirichlet_rows():
    x, cells = N3.unit_cube_tets(2)
    rng = np.random.default_rng(0)
    interior = (np.abs(x - 0.5) < 0.5 - 1e-12).all(axis=1)
    x[interior] += 0.05 * rng.standard_normal((int(interior.sum()), 3))
    rules = {k: S.tet_gauss_jacobi(4) for k in ("Fu", "Fp", "uu", "up", "pu", "pp")}
    prob = N3.Problem3D(x=x, cells=cells, dt=0.02, rho=1.06, mu=0.035, f=np.array([0.1, -0.3, 0.2]), rules=rules)
    n = prob.n
    xk = rng.standard_normal(4 * n)
    un = rng.standard_normal(3 * n)
    A = N3.assemble_J_raw(prob, xk, un).toarray()
    J = np.zeros_like(A)
    for j in range(4 * n):
        xc = xk.astype(complex)
        xc[j] += 1e-30j
        J[:, j] = N3.assemble_F_raw(prob, xc, un).imag / 1e-30
    assert np.abs(A - J).max() < 1e-13 * np.abs(A).max()
    # constant pressure is in the kernel of the cell integrals' pressure columns up to the boundary term:
    # sum of the momentum rows against a constant p vanishes in the interior (divergence theorem)
    v = np.zeros(4 * n)
    v[3 * n:] = 1.0
    r = (A @ v)[:3 * n].reshape(-1, 3)
    assert np.abs(r[interior]).max() < 1e-12 * np.abs(A).max()
