import sys, time, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import contextlib
import numpy as np, torch
kind = sys.argv[1]; res = float(sys.argv[2]); steps = int(sys.argv[3]); dt = float(sys.argv[4])
kw = {}
for a in sys.argv[5:]:
    k, v = a.split('='); kw[k] = eval(v)
t0 = time.time()
with contextlib.redirect_stdout(sys.stderr):
    if kind == "pressure":
        from cfd_hemodynamic_b200.src.scenarios.stenosis_pressure_structured import StenosisPressureStructuredSimulation as S
        sc = S("stabilized_schur_pressure_backflow", dt, 1.0, grade="severe", p_inlet=kw.pop("p_inlet", 80.0),
               R_resistance=kw.pop("R_resistance", 10.0), res=res, cell_type=kw.pop("cell_type", "triangle"), **kw)
    else:
        from cfd_hemodynamic_b200.src.scenarios.stenosis_mesh_variable import StenosisMeshVariableSimulation as S
        sc = S("stabilized_schur_backflow", dt, 1.0, grade="severe", v_max=kw.pop("v_max", 100.0), res=res, **kw)
s = sc.solver
torch.cuda.synchronize()
print(f"{kind}: cells {s._cells_host.shape[0]} dofs {s.N} setup {time.time()-t0:.1f}s nullspace {s._nullspace}")
for i in range(steps):
    torch.cuda.synchronize(); t1 = time.time()
    s.step_device()
    torch.cuda.synchronize()
    x = s.d_x
    print(f"step {i}: {1e3*(time.time()-t1):.1f} ms newton {s.its_snes} ksp {s.its_ksp} reason {s.reason} |u|max {float(x[:2*s.n].abs().max()):.3g} p range [{float(x[2*s.n:].min()):.4g},{float(x[2*s.n:].max()):.4g}]" + (f" p_c {s._p_c:.4g}" if hasattr(s, '_p_c') else ""))
