import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cfd_hemodynamic_b200.src.scenarios.lid_driven2D import LidDriven2DSimulation
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 224
TIGHT = dict(snes_rtol=1e-12, snes_stol=0.0, ksp_rtol=1e-11, ksp_restart=120)
sc = LidDriven2DSimulation("stabilized_schur", 0.01, 1.0, rho=1.0, mu=0.01, nx=nx, verbose=True, **TIGHT)
s = sc.solver
for k in range(3):
    try:
        s.solveStep()
    except Exception as e:
        print("FAILED step", k, repr(e), "last error:", s.hemo.lib.hemo_last_error(s.hemo._ctx))
        # look at the Krylov state
        break
    print("step", k, s.its_snes, s.its_ksp)
    s.u_prev.x.array[:] = s.u_sol.x.array[:]
    s.p_prev.x.array[:] = s.p_sol.x.array[:]
