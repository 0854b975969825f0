"""Lid-driven cavity Re = 100 marched to steady state; centre-line u(0.5, y) against Ghia, Ghia & Shin (1982), Table I
(the data set the reference ships as src/benchmark_data/lid_driven2D/plot_u_y_Ghia100.csv; its consumer is commented
out, src/scenarios/lid_driven2D.py:91-124)."""
import os, sys, time, json, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cfd_hemodynamic_b200.src.scenarios.lid_driven2D import LidDriven2DSimulation

GHIA_RE100 = np.array([[1.0000, 1.00000], [0.9766, 0.84123], [0.9688, 0.78871], [0.9609, 0.73722], [0.9531, 0.68717],
                       [0.8516, 0.23151], [0.7344, 0.00332], [0.6172, -0.13641], [0.5000, -0.20581], [0.4531, -0.21090],
                       [0.2813, -0.15662], [0.1719, -0.10150], [0.1016, -0.06434], [0.0703, -0.04775], [0.0625, -0.04192],
                       [0.0547, -0.03717], [0.0000, 0.00000]])


def centre_line_u(solver, ys, x0=0.5):
    """P1 interpolation of u_x at (x0, y) on the structured unit-square mesh."""
    mesh = solver.mesh
    x = mesh.geometry.x[:, :2]
    cells = mesh.geometry.dofmap
    U = solver.u_sol.x.array.reshape(-1, 2)
    X = x[cells]
    T = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]], axis=2)
    Tinv = np.linalg.inv(T)
    out = []
    for y in ys:
        lam = np.einsum("eij,ej->ei", Tinv, np.array([x0, y]) - X[:, 0])
        l0 = 1 - lam.sum(axis=1)
        c = np.nonzero((lam >= -1e-9).all(axis=1) & (l0 >= -1e-9))[0][0]
        w = np.array([l0[c], lam[c, 0], lam[c, 1]])
        out.append(float(w @ U[cells[c], 0]))
    return np.array(out)


if __name__ == "__main__":
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    dt = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
    T = float(sys.argv[3]) if len(sys.argv) > 3 else 30.0
    with contextlib.redirect_stdout(sys.stderr):
        sc = LidDriven2DSimulation("stabilized_schur", dt, T, rho=1.0, mu=0.01, nx=nx)
    s = sc.solver
    t0 = time.time()
    for i in range(int(round(T / dt))):
        s.solveStep()
        du = np.abs(s.u_sol.x.array - s.u_prev.x.array).max() / dt
        s.u_prev.x.array[:] = s.u_sol.x.array[:]
        s.p_prev.x.array[:] = s.p_sol.x.array[:]
    u = centre_line_u(s, GHIA_RE100[:, 0])
    err = np.abs(u - GHIA_RE100[:, 1])
    print(json.dumps({"nx": nx, "dt": dt, "T": T, "du_dt_final": du, "max_abs_err": float(err.max()), "u": u.tolist(),
                      "wall_s": time.time() - t0}))
