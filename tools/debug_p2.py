import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cfd_hemodynamic_b200.src.scenarios.stenosis_pressure_structured import StenosisPressureStructuredSimulation
from oracle.workload import CpuMarcher
from oracle import ns_oracle as O
TIGHT = dict(snes_rtol=1e-12, snes_atol=1e-8, snes_stol=0.0, ksp_rtol=1e-10, ksp_atol=1e-13, ksp_restart=150)
kw = dict(TIGHT); kw.update(eval(sys.argv[1]) if len(sys.argv) > 1 else {})
sc = StenosisPressureStructuredSimulation("stabilized_schur_pressure_backflow", 0.005, 0.02, grade="moderate", cell_type="triangle", p_inlet=2.0, R_resistance=float(os.environ.get("RR", "50")), res=0.8, L=20.0, x_position_stenosis=8.0, p_grade=2, verbose=True, **kw)
sc.setup()
s = sc.solver
m = CpuMarcher(sc, solver="lu", rtol=1e-12, atol=1e-8, stol=0.0)
n = s.n
for k in range(3):
    x0 = O.remove_nullspace(m.prob, m.x)
    F0 = O.assemble_F(m.prob, x0, m.un)
    print("oracle step", k, "|F0| =", np.linalg.norm(F0), "pconst", m.outlet["fs"].pconst, flush=True)
    try:
        m.step()
        print("oracle step", k, "ok", flush=True)
    except Exception as e:
        print("oracle FAILED", k, e, flush=True); break
    try:
        s.solveStep()
    except Exception as e:
        print("FAILED step", k, e, s.hemo.lib.hemo_last_error(s.hemo._ctx)); break
    print("step", k, s.its_snes, s.its_ksp, "pc", s._p_c, m.outlet["pc"], "du", np.abs(s.u_sol.x.array - m.x[:2*n]).max(), flush=True)
    s.u_prev.x.array[:] = s.u_sol.x.array[:]; s.p_prev.x.array[:] = s.p_sol.x.array[:]
