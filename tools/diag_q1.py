"""Diagnostic: lid cavity on Q1 quadrilaterals, per-step Newton / FGMRES counts."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cfd_hemodynamic_b200.src.scenarios.lid_driven2D import LidDriven2DSimulation
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ct = sys.argv[2] if len(sys.argv) > 2 else "quadrilateral"
kw = {}
for a in sys.argv[3:]:
    k, v = a.split("="); kw[k] = eval(v)
sc = LidDriven2DSimulation("stabilized_schur", 0.01, 0.03, rho=1, mu=0.01, nx=nx, cell_type=ct, verbose=True, **kw)
s = sc.solver
print("levels u", [ (d["P"].shape) for d in s.linear.levels[0]], "p", [(d["P"].shape) for d in s.linear.levels[1]])
for i in range(2):
    try:
        s.solveStep()
    except RuntimeError as e:
        print("step", i, "failed:", e, "its", s.its_snes, s.its_ksp)
        break
    print("step", i, "newton", s.its_snes, "ksp", s.its_ksp, "umax", np.abs(s.u_sol.x.array).max())
    s.u_prev.x.array[:] = s.u_sol.x.array[:]
