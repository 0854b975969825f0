"""The numpy oracles against a literal sympy transcription of the reference's UFL form
(oracle/form_mirror.py): residual of single cells, Jacobian by complex step of the mirrored
residual, and every boundary term — P1 triangles and non-affine Q1 quadrilaterals, mid-point and
BDF time schemes."""
import numpy as np
import pytest

from cfd_hemodynamic_b200.fem import mesh as M
from cfd_hemodynamic_b200.fem import quadrature as Q
from oracle import ns_oracle as O
from oracle import q1_oracle as Q1
from oracle.form_mirror import CellForms

X_TRI = np.array([[0.1, 0.05], [0.9, 0.2], [0.35, 0.8]])
X_QUAD = np.array([[0.0, 0.1], [1.0, 0.0], [0.15, 0.9], [1.2, 1.1]])          # tensor-ordered, non-affine


def _single_cell(cell_type, scheme, un_scale=1.0):
    X = X_TRI if cell_type == "triangle" else X_QUAD
    nv = X.shape[0]
    cells = np.arange(nv, dtype=np.int32)[None, :]
    mesh = M.Mesh(X, cells, cell_type="quadrilateral" if nv == 4 else None)
    if nv == 3:
        rules = {k: Q.triangle_gauss_jacobi(d) for k, d in dict(Fu=8, Fp=7, uu=8, up=7, pu=7, pp=6).items()}
    else:
        rules = {k: Q1.tensor_gauss(m) for k, m in dict(Fu=5, Fp=4, uu=5, up=4, pu=4, pp=3).items()}
    prob = O.Problem(x=X.copy(), cells=cells, h=mesh.h(2, np.arange(1)), dt=0.02, rho=1.06, mu=0.035,
                     f=np.array([0.1, -0.3]), rules=rules, facet_rule=Q.interval_gauss(3))
    rng = np.random.default_rng(3)
    U, P, Un = rng.standard_normal((nv, 2)), rng.standard_normal(nv), un_scale * rng.standard_normal((nv, 2))
    Uh = None
    if scheme == "bdf2":
        prob.theta, prob.a0 = 1.0, 1.5
        Uh = 2.0 * Un - 0.5 * (Un + 0.1 * rng.standard_normal((nv, 2)))
        prob.uh = Uh.reshape(-1)
    cf = CellForms(X, Un, float(prob.h[0]), prob.dt, prob.rho, prob.mu, prob.f, prob.eps0, theta=prob.theta,
                   a0=prob.a0, Uh=Uh)
    return mesh, prob, cf, U, P, Un


@pytest.mark.parametrize("scheme", ["midpoint", "bdf2"])
@pytest.mark.parametrize("cell_type", ["triangle", "quadrilateral"])
def test_cell_residual_and_jacobian_match_form_text(cell_type, scheme):
    mesh, prob, cf, U, P, Un = _single_cell(cell_type, scheme)
    nv = cf.nv
    K = O._kernels(prob)
    Uh = O._gather_history(prob)
    Fu, _ = K.element_F(prob, U[None], P[None], Un[None], prob.rules["Fu"], Uh)
    _, Fp = K.element_F(prob, U[None], P[None], Un[None], prob.rules["Fp"], Uh)
    Fu_m, Fp_m = cf.cell_residual(U, P, prob.rules["Fu"], prob.rules["Fp"])
    scale = max(np.abs(Fu).max(), np.abs(Fp).max())
    assert np.abs(Fu[0] - Fu_m).max() < 1e-12 * scale
    assert np.abs(Fp[0] - Fp_m).max() < 1e-12 * scale
    # Jacobian blocks: complex step of the mirrored residual, block by block with the block's rule
    Ae = O.element_matrices(prob, U.reshape(-1), P, Un.reshape(-1))[0]
    x0 = np.concatenate([U.reshape(-1), P])
    J = np.zeros((3 * nv, 3 * nv))
    for j in range(3 * nv):
        xc = x0.astype(complex)
        xc[j] += 1e-30j
        Uc, Pc = xc[:2 * nv].reshape(nv, 2), xc[2 * nv:]
        is_u = j < 2 * nv
        fu, _ = cf.cell_residual(Uc, Pc, prob.rules["uu" if is_u else "up"], prob.rules["Fp"])
        _, fp = cf.cell_residual(Uc, Pc, prob.rules["Fu"], prob.rules["pu" if is_u else "pp"])
        J[:2 * nv, j] = fu.reshape(-1).imag / 1e-30
        J[2 * nv:, j] = fp.imag / 1e-30
    assert np.abs(Ae - J).max() < 1e-11 * np.abs(Ae).max()


def test_zero_previous_velocity_branch_of_form_text():
    """u_prev = 0: the conditional of :101-103 takes the eps branch, tau_lsic vanishes."""
    mesh, prob, cf, U, P, Un = _single_cell("quadrilateral", "midpoint", un_scale=0.0)
    K = O._kernels(prob)
    Fu, _ = K.element_F(prob, U[None], P[None], Un[None], prob.rules["Fu"])
    Fu_m, _ = cf.cell_residual(U, P, prob.rules["Fu"], prob.rules["Fp"])
    assert np.abs(Fu[0] - Fu_m).max() < 1e-12 * np.abs(Fu).max()


@pytest.mark.parametrize("cell_type", ["triangle", "quadrilateral"])
def test_facet_terms_match_form_text(cell_type):
    mesh, prob, cf, U, P, Un = _single_cell(cell_type, "midpoint")
    nv = cf.nv
    K = O._kernels(prob)
    coef = dict(a_p=1.0, pconst=2.5, a_g=1.0, a_s=0.7, a_n=1.3, beta_n=100.0, a_b=0.9, beta_b=0.2)
    ext = M.exterior_facet_indices(mesh.topology)
    pairs = mesh.topology.facet_cell_pairs(ext)
    assert len(pairs) == nv
    X = prob.x
    for pair in pairs:
        fs = O.FacetSet(pairs=pair[None, :], **coef)
        F = K.facet_F(prob, fs, U[None], P[None], Un[None])[0]
        lf = int(pair[1])
        if nv == 3:
            va, vb = [(1, 2), (0, 2), (0, 1)][lf]
            inside = X[lf]
        else:
            va, vb = M.QUAD_FACETS[lf]
            inside = X.mean(axis=0)
        t = X[vb] - X[va]
        length = np.linalg.norm(t)
        nrm = np.array([t[1], -t[0]]) / length
        if nrm @ (0.5 * (X[va] + X[vb]) - inside) < 0:
            nrm = -nrm
        Fm = cf.facet_residual(U, P, (va, vb), nrm, length, prob.facet_rule, **coef)
        assert np.abs(F - Fm).max() < 1e-12 * np.abs(F).max(), (lf, np.abs(F - Fm).max())
