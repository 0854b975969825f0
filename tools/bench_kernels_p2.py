"""Micro-benchmark of the P2-P2 triangle assembly kernels (CUDA events, L2 flushed between reps):
    python tools/bench_kernels_p2.py [nx, default 150 -> 45 k perturbed cells]
Jacobian / residual assembly (cells at the reference degrees 20 / 18 / 18 / 16 + facet terms + Dirichlet rows) and the SpMV on the
P2 node graph.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cfd_hemodynamic_b200._lib import Hemo
from cfd_hemodynamic_b200.fem import discretization as D, mesh as M, quadrature as Q
from oracle import ns_oracle as O
from oracle import pk_oracle as PK
from tests import common as T

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 150
reps = 5
t0 = time.time()
mesh = T.perturbed_square(nx, nx, seed=0, amp=0.2)
x, cells6 = D.p2_nodes(mesh)
deg = {"Fu": 20, "Fp": 18, "uu": 20, "up": 18, "pu": 18, "pp": 16}
prob = O.Problem(x=x, cells=cells6, h=PK.cell_diameter(x, cells6), dt=0.01, rho=1.0, mu=0.01, f=np.zeros(2),
                 rules={k: Q.triangle_rule(d) for k, d in deg.items()}, facet_rule=Q.interval_gauss(4))
ext = M.exterior_facet_indices(mesh.topology)
nv = mesh.geometry.x.shape[0]
wall_nodes = np.union1d(np.unique(mesh.topology.facet_vertices[ext]), nv + ext)
n = prob.n
h = Hemo(0)
g, _ = T.setup_gpu(h, mesh, prob, [(ext, dict(a_p=1.0, a_g=1.0))], [("u", wall_nodes, np.zeros(2 * n))])
torch.cuda.synchronize()
print("setup s", time.time() - t0, "cells", cells6.shape[0], "nodes", n, "nnz", h.nnz, file=sys.stderr)
dev = h.device
u = np.stack([np.sin(2.1 * x[:, 0] + 0.3) * np.cos(1.7 * x[:, 1]), -np.cos(1.3 * x[:, 0]) * np.sin(2.3 * x[:, 1] + 0.2)], axis=1)
xd = torch.tensor(np.concatenate([u.reshape(-1), np.sin(x[:, 0])]), device=dev)
und = torch.tensor(0.9 * u.reshape(-1), device=dev)
vals = torch.zeros(h.nnz, dtype=torch.float64, device=dev)
b = torch.zeros(3 * n, dtype=torch.float64, device=dev)
y = torch.zeros_like(b)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)


def timeit(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


E = cells6.shape[0]
out = dict(cells=E, nodes=n, nnz=h.nnz, cell_type="triangle_p2", rule_points={k: len(prob.rules[k][1]) for k in deg})
out["jacobian_ms"] = timeit(lambda: h.assemble_jacobian(xd, und, vals))
out["residual_ms"] = timeit(lambda: h.assemble_residual(xd, und, g, b))
out["spmv_ms"] = timeit(lambda: h.spmv(vals, xd, y))
out["jac_Mcells_s"] = E / out["jacobian_ms"] / 1e3
out["res_Mcells_s"] = E / out["residual_ms"] / 1e3
out["jac_nnz_per_s"] = h.nnz / out["jacobian_ms"] * 1e3
print(json.dumps(out))
