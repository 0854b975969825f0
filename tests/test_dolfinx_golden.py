"""Consumes reference-pinned golden files written by tools/dump_dolfinx_golden.py inside a real DOLFINx v0.9 container
(`tests/golden/dolfinx_*.npz`): geometry and dofmaps as DOLFINx numbers them, Basix quadrature rules, the assembled
Jacobian (A.getValuesCSR) and residual at a seeded state, and the marched solution.  Skipped while no such file exists —
DOLFINx cannot be installed in this repository's container, which is why the oracle is "parity unpinned" (DESIGN.md §2);
the day a file is dropped in, these tests pin it (and the CUDA path under `-m gpu`)."""
import glob
import os

import numpy as np
import pytest

FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "dolfinx_*.npz")))
pytestmark = pytest.mark.skipif(not FILES, reason="no tests/golden/dolfinx_*.npz (run tools/dump_dolfinx_golden.py in the "
                                                  "reference's DOLFINx container)")


def problem_from_golden(g):
    """Oracle Problem in DOLFINx's dof numbering with the Basix rules of the file."""
    from oracle import ns_oracle as O
    gdim = 2
    cells = g["V_dofmap"].astype(np.int32)
    n = int(cells.max()) + 1
    x = g["dof_coordinates"][:n, :gdim]
    rules = {k: (g[f"rule_{k}_pts"], g[f"rule_{k}_wts"]) for k in ("Fu", "Fp", "uu", "up", "pu", "pp")}
    prob = O.Problem(x=x.copy(), cells=cells, h=g["h"].astype(float), dt=float(g["dt"]), rho=float(g["rho"]),
                     mu=float(g["mu"]), f=np.asarray(g["f"], dtype=float).reshape(-1)[:2], rules=rules,
                     facet_rule=(g["frule_pts"], g["frule_wts"]))
    prob.facet_sets = [O.FacetSet(pairs=g["ext_pairs"].astype(np.int32), a_p=1.0, a_g=1.0)]
    bcs = []
    for i in range(int(g["n_bcs"])):
        block = str(g[f"bc{i}_block"])
        dofs = g[f"bc{i}_dofs"].astype(np.int64)
        bcs.append((block, dofs, g[f"bc{i}_values"].astype(float)))
    prob.bcs = bcs
    return prob


@pytest.mark.parametrize("path", FILES)
def test_oracle_matches_dolfinx(path):
    import scipy.sparse as sp
    from oracle import ns_oracle as O
    g = np.load(path)
    assert np.array_equal(g["V_dofmap"], g["Q_dofmap"]), "V and Q numbered differently: extend the test with the Q map"
    prob = problem_from_golden(g)
    n = prob.n
    rp, ci = O.sparsity_pattern(prob)
    assert np.array_equal(rp, g["A_indptr"]) and np.array_equal(ci, g["A_indices"])        # pattern bit-exact
    u, p, un = g["u"][:2 * n], g["p"][:n], g["un"][:2 * n]
    A = O.assemble_J(prob, u, p, un)
    A_ref = sp.csr_matrix((g["A_data"], g["A_indices"], g["A_indptr"]), shape=A.shape)
    diff = (A - A_ref).tocsr()
    assert (np.linalg.norm(diff.data) if diff.nnz else 0.0) <= 1e-12 * np.linalg.norm(A_ref.data)
    b = O.assemble_F(prob, np.concatenate([u, p]), un)
    assert np.linalg.norm(b - g["b"]) <= 1e-12 * np.linalg.norm(g["b"])
    xk, unk = np.zeros(3 * n), np.zeros(2 * n)
    for _ in range(int(g["n_steps"])):
        xk = O.remove_nullspace(prob, xk)
        xk, its, reason = O.newton_solve(prob, xk, unk, rtol=1e-12, stol=0.0)
        assert reason > 0
        unk = xk[:2 * n].copy()
    assert np.linalg.norm(xk[:2 * n] - g["u_end"]) <= 1e-8 * np.linalg.norm(g["u_end"])
    pe, pr = xk[2 * n:] - xk[2 * n:].mean(), g["p_end"] - g["p_end"].mean()
    assert np.linalg.norm(pe - pr) <= 1e-8 * np.linalg.norm(pr)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES)
def test_cuda_path_matches_dolfinx(path):
    torch = pytest.importorskip("torch")
    from cfd_hemodynamic_b200._lib import Hemo
    from cfd_hemodynamic_b200.fem import discretization as D
    from cfd_hemodynamic_b200.fem.mesh import Mesh
    from tests import common as T
    g = np.load(path)
    prob = problem_from_golden(g)
    n = prob.n
    mesh = Mesh(prob.x, prob.cells)
    h = Hemo(0)
    dev = h.device
    h.set_mesh(torch.tensor(prob.x, device=dev), torch.tensor(prob.cells, device=dev), torch.tensor(prob.h, device=dev))
    nrowptr, ncol = D.node_graph(prob.cells, n)
    h.set_node_graph(torch.tensor(nrowptr, device=dev), torch.tensor(ncol, device=dev))
    for k, bid in T.BLOCK_ID.items():
        h.set_quadrature(bid, *prob.rules[k])
    h.set_facet_quadrature(*prob.facet_rule)
    h.set_params(prob.dt, prob.rho, prob.mu, prob.f, prob.eps0)
    fc, fm = D.pairs_by_cell(prob.facet_sets[0].pairs)
    h.set_facet_set(0, torch.tensor(fc, device=dev), torch.tensor(fm, device=dev), a_p=1.0, a_g=1.0)
    bcs = [(b, np.unique(d // 2) if b == "u" else d, v) for b, d, v in prob.bcs]
    flag, mult, cellflag, gv = D.dirichlet_arrays(n, prob.cells, bcs)
    h.set_bc(torch.tensor(flag, device=dev), torch.tensor(mult, device=dev), torch.tensor(cellflag, device=dev))
    xd = torch.tensor(np.concatenate([g["u"][:2 * n], g["p"][:n]]), device=dev)
    und = torch.tensor(g["un"][:2 * n], device=dev)
    vals = torch.zeros(h.nnz, dtype=torch.float64, device=dev)
    b = torch.zeros(3 * n, dtype=torch.float64, device=dev)
    h.assemble_jacobian(xd, und, vals)
    h.assemble_residual(xd, und, torch.tensor(gv, device=dev), b)
    rowptr, col = h.get_pattern()
    assert np.array_equal(rowptr.cpu().numpy(), g["A_indptr"]) and np.array_equal(col.cpu().numpy(), g["A_indices"])
    assert np.linalg.norm(vals.cpu().numpy() - g["A_data"]) <= 1e-12 * np.linalg.norm(g["A_data"])
    assert np.linalg.norm(b.cpu().numpy() - g["b"]) <= 1e-12 * np.linalg.norm(g["b"])
    h.close()


def _write_synthetic(path, seed=2):
    """A file in the dump format, produced by the ORACLE on a permuted numbering: keeps the reader above honest while
    no real DOLFINx file exists (it pins nothing — the values come from the code under test)."""
    import scipy.sparse as sp
    from cfd_hemodynamic_b200.fem import mesh as M
    from oracle import ns_oracle as O
    from tests import common as T
    from tools.dump_dolfinx_golden import smooth_fields
    m = M.create_unit_square(None, 5, 4)
    n = m.geometry.x.shape[0]
    perm = np.random.default_rng(seed).permutation(n)
    x = np.empty((n, 2))
    x[perm] = m.geometry.x[:, :2]
    cells = perm[m.geometry.dofmap].astype(np.int32)
    pm = M.Mesh(x, cells)
    prob = T.make_problem(pm, dt=0.01, rho=1.0, mu=0.01, f=(0.0, 0.0))
    ext = M.exterior_facet_indices(pm.topology)
    prob.facet_sets = [O.FacetSet(pairs=pm.topology.facet_cell_pairs(ext), a_p=1.0, a_g=1.0)]
    walls = np.nonzero(np.isclose(x[:, 0], 0) | np.isclose(x[:, 0], 1) | np.isclose(x[:, 1], 0))[0]
    lid = np.nonzero(np.isclose(x[:, 1], 1.0) & (x[:, 0] > 1e-10) & (x[:, 0] < 1 - 1e-10))[0]
    g1 = np.zeros(2 * n)
    g1[0::2] = 1.0
    prob.bcs = T.oracle_bcs(prob, [("u", walls, np.zeros(2 * n)), ("u", lid, g1)])
    out = dict(geometry_x=m.geometry.x, geometry_dofmap=m.geometry.dofmap, cell_name=np.array("triangle"), V_dofmap=cells,
               Q_dofmap=cells, V_bs=np.array(2), dof_coordinates=np.hstack([x, np.zeros((n, 1))]), h=prob.h,
               ext_pairs=prob.facet_sets[0].pairs, n_bcs=np.array(2), dt=np.array(prob.dt), rho=np.array(prob.rho),
               mu=np.array(prob.mu), f=prob.f, frule_pts=prob.facet_rule[0], frule_wts=prob.facet_rule[1])
    for i, (b, d, v) in enumerate(prob.bcs):
        out[f"bc{i}_block"], out[f"bc{i}_dofs"], out[f"bc{i}_values"] = np.array(b), d, v
    for k, (pts, wts) in prob.rules.items():
        out[f"rule_{k}_pts"], out[f"rule_{k}_wts"] = pts, wts
    u, p, un = smooth_fields(x)
    A = O.assemble_J(prob, u, p, un)
    rp, ci = O.sparsity_pattern(prob)                 # DOLFINx keeps the full pattern (explicit zeros on Dirichlet rows)
    N = 3 * n
    keys_pat = np.repeat(np.arange(N, dtype=np.int64), np.diff(rp)) * N + ci
    Ac = A.tocoo()
    data = np.zeros(ci.shape[0])
    data[np.searchsorted(keys_pat, Ac.row.astype(np.int64) * N + Ac.col)] = Ac.data
    A = sp.csr_matrix((data, ci, rp), shape=(N, N))
    out.update(u=u, p=p, un=un, A_indptr=A.indptr, A_indices=A.indices, A_data=A.data,
               b=O.assemble_F(prob, np.concatenate([u, p]), un))
    xk, unk = np.zeros(3 * n), np.zeros(2 * n)
    for _ in range(2):
        xk = O.remove_nullspace(prob, xk)
        xk, _, reason = O.newton_solve(prob, xk, unk, rtol=1e-12, stol=0.0)
        unk = xk[:2 * n].copy()
    out.update(u_end=xk[:2 * n], p_end=xk[2 * n:], n_steps=np.array(2))
    np.savez_compressed(path, **out)
