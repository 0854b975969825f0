"""Parameter sweep of the block preconditioner on the pressure-driven stenosis workload (north-star scenario):
iterations and time per step for a few Schur / Krylov settings.  Usage: sweep_stenosis.py <res> <steps> <warmup>"""
import os, sys, time, json, contextlib, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cfd_hemodynamic_b200.src.scenarios.stenosis_pressure_structured import StenosisPressureStructuredSimulation

res = float(sys.argv[1]) if len(sys.argv) > 1 else 0.015
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
W = int(sys.argv[3]) if len(sys.argv) > 3 else 2
only = sys.argv[4].split(";") if len(sys.argv) > 4 else None      # configuration names separated by ";"
CONFIGS = {
    "base(selfp,restart60,step)": dict(),
    "restart20": dict(ksp_restart=20),
    "restart30": dict(ksp_restart=30),
    "restart40": dict(ksp_restart=40),
    "restart120": dict(ksp_restart=120),
    "restart200": dict(ksp_restart=200),
    "pc_newton": dict(pc_rebuild="newton"),
    "restart200+pc_newton": dict(ksp_restart=200, pc_rebuild="newton"),
    "cycles_p2": dict(amg_cycles_p=2),
    "cheb3": dict(cheb_degree=3),
    "cycles_u2": dict(amg_cycles_u=2),
    "restart200+cycles_p2+u2": dict(ksp_restart=200, amg_cycles_p=2, amg_cycles_u=2),
    "laplace": dict(schur_mode="laplace"),
}
for name, kw in CONFIGS.items():
    if only and name not in only:
        continue
    t0 = time.time()
    try:
        with contextlib.redirect_stdout(sys.stderr):
            sc = StenosisPressureStructuredSimulation("stabilized_schur_pressure_backflow", 1e-3, 1.0, grade="severe", p_inlet=80.0,
                                                      R_resistance=10.0, res=res, cell_type="triangle", **kw)
        s = sc.solver
        for _ in range(W):
            s.step_device()
        torch.cuda.synchronize()
        t1 = time.time()
        newton = ksp = 0
        for _ in range(K):
            s.step_device()
            newton += s.its_snes
            ksp += s.its_ksp
        torch.cuda.synchronize()
        dt = (time.time() - t1) / K
        print(json.dumps({"config": name, "cells": int(s._cells_host.shape[0]), "ms_per_step": 1e3 * dt, "newton_per_step": newton / K,
                          "fgmres_per_step": ksp / K, "fgmres_per_solve": ksp / max(newton, 1), "MDOF_steps_per_s": s.N / dt / 1e6,
                          "build_s": t1 - t0}), flush=True)
        s.hemo.close()
    except Exception as e:
        print(json.dumps({"config": name, "error": repr(e)[:200]}), flush=True)
    del sc, s
    gc.collect()
    torch.cuda.empty_cache()
