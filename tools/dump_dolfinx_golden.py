#!/usr/bin/env python
"""Dump reference-pinned golden data from a REAL DOLFINx v0.9 run of the reference.

Runs inside the reference's container (singularity.def:2 / Dockerfile:1, `dolfinx/dolfinx:v0.9.0`), from the root of
the reference checkout:

    python /path/to/this/repo/tools/dump_dolfinx_golden.py --case lid --nx 8 --out dolfinx_lid_nx8.npz
    python /path/to/this/repo/tools/dump_dolfinx_golden.py --case dfg --out dolfinx_dfg.npz

and the resulting file goes to `tests/golden/` of this repository; `tests/test_dolfinx_golden.py` consumes every
`tests/golden/dolfinx_*.npz` it finds (oracle on CPU, CUDA path under `-m gpu`) and is skipped while none exists.
Nothing here can run in this repository's own container (no DOLFINx): the script is written against the API the
reference itself uses (src/solvers/stabilized_schur.py, src/scenarios/*.py) and is deliberately small.

What is written (all arrays in DOLFINx's own numbering; serial run):
    geometry_x, geometry_dofmap, cell_name
    V_dofmap, Q_dofmap, V_bs, dof_coordinates       (dof -> coordinates; the library works in dof numbering)
    h                                                mesh.h(tdim, all cells)
    ext_pairs                                        (cell, local facet) of every exterior facet
    tag_<id>_pairs                                   the same for every facet tag the scenario defines
    bc<i>_block ('u'|'p'), bc<i>_dofs (unrolled), bc<i>_values (full g.x.array)   — in list order
    rule_<block>_pts / _wts, degree_<block>         Basix default rule for the UFL-estimated degree of each block form
    frule_pts / frule_wts                            facet rule
    dt, rho, mu, f
    u, p, un                                         seeded smooth state the operators are evaluated at
    A_indptr, A_indices, A_data                      A.getValuesCSR() after assembleJacobian          (1e-12 target)
    b                                                F after assembleResidual (lifting + set_bc)      (1e-12 target)
    u_end, p_end, n_steps                            state after n_steps solveStep() with tight tolerances (1e-8 target)
"""
import argparse
import sys

import numpy as np


def smooth_fields(x, seed=1, U=1.0):
    """Same seeded fields as tests/common.py:smooth_fields, evaluated at dof coordinates."""
    rng = np.random.default_rng(seed)
    n = x.shape[0]
    u = np.empty((n, 2))
    u[:, 0] = U * np.sin(2.1 * x[:, 0] + 0.3) * np.cos(1.7 * x[:, 1])
    u[:, 1] = -U * np.cos(1.3 * x[:, 0]) * np.sin(2.3 * x[:, 1] + 0.2)
    un = 0.9 * u + 0.05 * U * np.cos(3.0 * x[:, :1] + x[:, 1:])
    u += 0.01 * U * rng.standard_normal(u.shape)
    un += 0.01 * U * rng.standard_normal(u.shape)
    p = np.sin(1.1 * x[:, 0]) * x[:, 1] + 0.01 * rng.standard_normal(n)
    return u.reshape(-1), p, un.reshape(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="lid", choices=["lid", "dfg"])
    ap.add_argument("--nx", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--dt", type=float, default=0.01)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()

    import basix
    import ufl
    from dolfinx import fem
    from dolfinx.fem import IntegralType, compute_integration_domains
    from dolfinx.mesh import exterior_facet_indices
    from petsc4py import PETSc
    from ufl.algorithms import estimate_total_polynomial_degree, expand_derivatives

    # tight tolerances for the marched solution; read through setFromOptions()
    opts = PETSc.Options()
    opts["nonlinear_snes_rtol"] = 1e-12
    opts["nonlinear_snes_stol"] = 0.0
    opts["nonlinear_ksp_rtol"] = 1e-11

    sys.path.insert(0, ".")
    if args.case == "lid":
        from src.scenarios.lid_driven2D import LidDriven2DSimulation as Scn
        sc = Scn("stabilized_schur", args.dt, args.dt * args.steps, nx=args.nx, mu=0.01)
    else:
        from src.scenarios.dfg_1 import DFG1Benchmark as Scn        # class name as in src/scenarios/dfg_1.py
        sc = Scn("stabilized_schur", args.dt, args.dt * args.steps)
    s = sc.solver                      # the scenario constructor has already called setup() once
    mesh = s.mesh
    tdim = mesh.topology.dim
    fdim = tdim - 1
    mesh.topology.create_connectivity(fdim, tdim)
    ncells = mesh.topology.index_map(tdim).size_local
    out = {}
    out["geometry_x"] = mesh.geometry.x.copy()
    out["geometry_dofmap"] = np.asarray(mesh.geometry.dofmap).reshape(ncells, -1)
    out["cell_name"] = np.array(mesh.topology.cell_name())
    out["V_dofmap"] = np.asarray(s.V.dofmap.list).reshape(ncells, -1)
    out["Q_dofmap"] = np.asarray(s.Q.dofmap.list).reshape(ncells, -1)
    out["V_bs"] = np.array(s.V.dofmap.index_map_bs)
    out["dof_coordinates"] = s.V.tabulate_dof_coordinates().copy()
    out["h"] = mesh.h(tdim, np.arange(ncells, dtype=np.int32))
    ext = exterior_facet_indices(mesh.topology)
    out["ext_pairs"] = np.asarray(compute_integration_domains(IntegralType.exterior_facet, mesh.topology, ext, fdim)).reshape(-1, 2)
    ft = getattr(sc, "facet_tags", None)
    if ft is not None:
        for val in np.unique(ft.values):
            fac = ft.find(val)
            out[f"tag_{int(val)}_pairs"] = np.asarray(
                compute_integration_domains(IntegralType.exterior_facet, mesh.topology, fac, fdim)).reshape(-1, 2)
    bcs = [("u", bc) for bc in s.bcu_d] + [("p", bc) for bc in s.bcp_d]
    for i, (block, bc) in enumerate(bcs):
        out[f"bc{i}_block"] = np.array(block)
        out[f"bc{i}_dofs"] = np.asarray(bc.dof_indices()[0])
        out[f"bc{i}_values"] = bc.g.x.array.copy()
    out["n_bcs"] = np.array(len(bcs))
    # quadrature FFCx selects: estimated degree of each block form -> Basix default rule
    Fb = ufl.extract_blocks(s.F)
    du, dp = ufl.TrialFunctions(s.VQ)
    Jb = ufl.extract_blocks(ufl.derivative(s.F, (s.u_sol, s.p_sol), (du, dp)))
    forms = {"Fu": Fb[0], "Fp": Fb[1], "uu": Jb[0][0], "up": Jb[0][1], "pu": Jb[1][0], "pp": Jb[1][1]}
    ctype = getattr(basix.CellType, mesh.topology.cell_name())
    for k, frm in forms.items():
        dx_part = sum((itg for itg in frm.integrals() if itg.integral_type() == "cell"), ufl.form.Zero()) \
            if hasattr(ufl.form, "Zero") else frm
        degs = [estimate_total_polynomial_degree(expand_derivatives(ufl.Form([itg])))
                for itg in frm.integrals() if itg.integral_type() == "cell"]
        deg = int(max(degs))
        pts, wts = basix.make_quadrature(ctype, deg)
        out[f"degree_{k}"] = np.array(deg)
        out[f"rule_{k}_pts"], out[f"rule_{k}_wts"] = np.asarray(pts), np.asarray(wts)
    fdegs = [estimate_total_polynomial_degree(expand_derivatives(ufl.Form([itg])))
             for frm in forms.values() for itg in frm.integrals() if itg.integral_type() == "exterior_facet"]
    fpts, fwts = basix.make_quadrature(basix.CellType.interval, int(max(fdegs)) if fdegs else 2)
    out["frule_pts"], out["frule_wts"] = np.asarray(fpts).reshape(-1), np.asarray(fwts)
    out["dt"], out["rho"], out["mu"] = np.array(s.dt.value), np.array(s.rho.value), np.array(s.mu.value)
    out["f"] = np.asarray(s.f.value).copy()

    # ---- operators at a fixed smooth state --------------------------------------------------------------
    xd = out["dof_coordinates"][:, :2]
    u, p, un = smooth_fields(xd)
    nu = s.u_sol.x.array.shape[0]
    s.u_prev.x.array[:] = un[:nu]
    xvec = s.x_n.copy()
    xvec.array[:nu] = u[:nu]
    xvec.array[nu:] = p[:s.p_sol.x.array.shape[0]]
    allbc = [*s.bcu_d, *s.bcp_d]
    bvec = s.b.duplicate()
    s.assembleResidual(None, xvec, bvec, bcs=allbc)          # also copies x into u_sol / p_sol
    s.assembleJacobian(None, xvec, s.A, None, bcs=allbc)
    ia, ja, va = s.A.getValuesCSR()
    out["u"], out["p"], out["un"] = u, p, un
    out["A_indptr"], out["A_indices"], out["A_data"] = np.asarray(ia), np.asarray(ja), np.asarray(va)
    out["b"] = bvec.array.copy()

    # ---- marched solution (zero initial state, the scenario's own time loop semantics) -------------------
    s.u_prev.x.array[:] = 0.0
    s.p_prev.x.array[:] = 0.0
    s.x_n.array[:] = 0.0
    for _ in range(args.steps):
        s.solveStep()
        s.u_prev.x.array[:] = s.u_sol.x.array[:]            # src/scenario.py:306-307
        s.p_prev.x.array[:] = s.p_sol.x.array[:]
    out["u_end"], out["p_end"], out["n_steps"] = s.u_sol.x.array.copy(), s.p_sol.x.array.copy(), np.array(args.steps)
    np.savez_compressed(args.out, **out)
    print(f"wrote {args.out}: {ncells} cells, {out['A_data'].shape[0]} nnz, quadrature degrees "
          + ", ".join(f"{k}={int(out['degree_' + k])}" for k in forms))


if __name__ == "__main__":
    main()
