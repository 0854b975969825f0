"""Micro-benchmark of the assembly / SpMV kernels (CUDA events, L2 flushed between reps)."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cfd_hemodynamic_b200._lib import Hemo
from cfd_hemodynamic_b200.fem import mesh as M, discretization as D
from tests import common as T

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 707
cell_type = sys.argv[2] if len(sys.argv) > 2 else "triangle"
reps = 10
t0 = time.time()
# perturbed interior vertices: non-affine quadrilaterals take the general path
mesh = T.perturbed_square(nx, nx, seed=0, amp=0.2, cell_type=cell_type)
prob = T.make_problem(mesh, dt=0.01, rho=1.0, mu=0.01, f=(0, 0))
ext = M.exterior_facet_indices(mesh.topology)
h = Hemo(0)
x = prob.x; n = prob.n
walls = np.nonzero(np.isclose(x[:, 0], 0) | np.isclose(x[:, 0], 1) | np.isclose(x[:, 1], 0) | np.isclose(x[:, 1], 1))[0]
bcs = [("u", walls, np.zeros(2 * n))]
g, _ = T.setup_gpu(h, mesh, prob, [(ext, dict(a_p=1.0, a_g=1.0))], bcs)
torch.cuda.synchronize()
print("setup s", time.time() - t0, "cells", prob.cells.shape[0], "nnz", h.nnz)
u, p, un = T.smooth_fields(x)
dev = h.device
xd = torch.tensor(np.concatenate([u, p]), device=dev); und = torch.tensor(un, device=dev)
vals = torch.zeros(h.nnz, dtype=torch.float64, device=dev)
b = torch.zeros(3 * n, dtype=torch.float64, device=dev); y = torch.zeros_like(b)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

def timeit(fn):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

E = prob.cells.shape[0]
out = {}
out["jacobian_ms"] = timeit(lambda: h.assemble_jacobian(xd, und, vals))
out["residual_ms"] = timeit(lambda: h.assemble_residual(xd, und, g, b))
out["spmv_ms"] = timeit(lambda: h.spmv(vals, xd, y))
out["dot_ms"] = timeit(lambda: h.dot(xd, y))
out["cells"] = E; out["nnz"] = h.nnz; out["cell_type"] = cell_type
out["res_Mcells_s"] = E / out["residual_ms"] / 1e3
out["jac_nnz_per_s"] = h.nnz / out["jacobian_ms"] * 1e3
if hasattr(h, "prof_enable"):
    h.prof_enable(True)
    for _ in range(5):
        h.assemble_jacobian(xd, und, vals); h.assemble_residual(xd, und, g, b)
    for cls, name in ((1, "cell_jacobian"), (2, "gather_matrix"), (3, "cell_residual")):
        ms, cnt = h.prof_get(cls)
        out[name + "_ms"] = ms / max(cnt, 1)
    h.prof_enable(False)
out["jac_Mcells_s"] = E / out["jacobian_ms"] / 1e3
out["spmv_GBs"] = (8 * h.nnz + 4 * h.nnz_node + 20 * 3 * n) / out["spmv_ms"] / 1e6
print(json.dumps(out))
