"""Micro-benchmark of the tetrahedron kernels (CUDA events, L2 flushed between reps):
    python tools/bench_kernels_tet.py [cubes per edge, default 55 -> 1.0 M tetrahedra] [degree of the J_uu / F_u rule, default 12]
Jacobian / residual assembly (cell + all-facet terms + Dirichlet rows on the boundary), SpMV on the 3-D CSR
layout and one application of the first 3-D preconditioner.  Prints one JSON line (profiles/r02_tet_kernels_280k.json)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from cfd_hemodynamic_b200._lib import Hemo
from cfd_hemodynamic_b200.fem import discretization as D
from cfd_hemodynamic_b200.fem import mesh as M
from cfd_hemodynamic_b200.fem import quadrature as Q
from cfd_hemodynamic_b200.linear_solver import BlockSchurSolver

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 55
deg = int(sys.argv[2]) if len(sys.argv) > 2 else 12
reps = 10
t0 = time.time()
mesh = M.create_unit_cube(None, nc, nc, nc)
x = np.ascontiguousarray(mesh.geometry.x)
cells = np.ascontiguousarray(mesh.geometry.dofmap, dtype=np.int32)
rng = np.random.default_rng(0)
interior = (np.abs(x - 0.5) < 0.5 - 1e-12).all(axis=1)
x[interior] += 0.2 / nc * (rng.random((int(interior.sum()), 3)) - 0.5)
n, E = x.shape[0], cells.shape[0]
hcell = mesh.h(3, np.arange(E))
hm = Hemo(0)
dev = hm.device
T = lambda a, dt=torch.float64: torch.tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
hm.set_mesh(T(x), T(cells, torch.int32), T(hcell))
nrowptr, ncol = D.node_graph(cells, n)
hm.set_node_graph(T(nrowptr, torch.int32), T(ncol, torch.int32))
for block, d in enumerate((deg, deg - 1, deg, deg - 1, deg - 1, deg - 2)):
    hm.set_quadrature(block, *Q.tetrahedron_rule(d))
hm.set_facet_quadrature(*Q.triangle_rule(2))
dt, rho, mu = 0.005, 1.0, 0.02
hm.set_params(dt, rho, mu, np.zeros(2), float(np.finfo(np.float64).resolution))
hm.set_body_force3(np.zeros(3))
ext = M.exterior_facet_indices(mesh.topology)
fc, fm = D.facet_set_by_cell(mesh, ext)
hm.set_facet_set(0, T(fc, torch.int32), T(fm, torch.int32), a_p=1.0, a_g=1.0)
boundary = np.nonzero(~interior)[0]
flag, mult, cellflag, g = D.dirichlet_arrays(n, cells, [("u", boundary, np.zeros(3 * n))], gdim=3)
hm.set_bc(T(flag, torch.uint8), T(mult), T(cellflag, torch.uint8))
torch.cuda.synchronize()
print("setup s", time.time() - t0, "cells", E, "nodes", n, "nnz", hm.nnz, file=sys.stderr)
u = np.stack([np.sin(2.1 * x[:, 0] + 0.3) * np.cos(1.7 * x[:, 1]), -np.cos(1.3 * x[:, 0]) * np.sin(2.3 * x[:, 2] + 0.2),
              0.5 * np.sin(x[:, 1] + x[:, 2])], axis=1)
u[boundary] = 0.0
un = 0.9 * u
xd, und, gd = T(np.concatenate([u.reshape(-1), np.sin(x[:, 0])])), T(un.reshape(-1)), T(g)
vals = torch.zeros(hm.nnz, dtype=torch.float64, device=dev)
b = torch.zeros(4 * n, dtype=torch.float64, device=dev)
y = torch.zeros_like(b)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)


def timeit(fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


out = dict(cells=E, nodes=n, nnz=hm.nnz, rule_points=len(Q.tetrahedron_rule(deg)[1]), cell_type="tetrahedron")
out["jacobian_ms"] = timeit(lambda: hm.assemble_jacobian(xd, und, vals))
out["residual_ms"] = timeit(lambda: hm.assemble_residual(xd, und, gd, b))
out["spmv_ms"] = timeit(lambda: hm.spmv(vals, xd, y))
out["jac_Mcells_s"] = E / out["jacobian_ms"] / 1e3
out["res_Mcells_s"] = E / out["residual_ms"] / 1e3
out["jac_nnz_per_s"] = hm.nnz / out["jacobian_ms"] * 1e3
# algorithmic bytes of the SpMV: 16 values + 1 column index per node pair, 4 row pointers... per node: x, y, rowptr
out["spmv_GBs"] = (8 * hm.nnz + 4 * hm.nnz_node + (8 * 4 * 2 + 4) * n) / out["spmv_ms"] / 1e6
ks = BlockSchurSolver(hm, nrowptr, ncol, boundary, np.zeros(0, np.int64), dt=dt, rho=rho, mu=mu, project_pressure=True)
ks.setup(vals)
out["pc_apply_ms"] = timeit(lambda: hm.pc_apply(vals, b, y))
hm.assemble_residual(xd, und, gd, b)
its, rel = ks.solve(vals, b, y)
out["fgmres_its"], out["fgmres_rel"] = its, rel
out["fgmres_ms"] = timeit(lambda: ks.solve(vals, b, y))
print(json.dumps(out))
