#!/usr/bin/env python
"""Benchmark of the stabilized_schur per-timestep hot path (BASELINE.json metric:
DOF-timesteps/s; assembly cells/s and nnz/s are reported beside it).

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port
                                                           # (DOLFINx/PETSc cannot be installed)

A "step" is one time step (`Solver.solveStep`) of the workload: Newton with
Jacobian + residual assemblies and a preconditioned FGMRES solve per
iteration.  `value` is measured with everything resident in HBM
(`Solver.step_device`), `e2e` through the reference-facing plugin API with host
buffers (`Solver.solveStep` + the host-side u_prev <- u_sol shift the
reference's time loop performs, src/scenario.py:306-307).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: lid_driven2D on a refined structured mesh (~1M cells)
    "lid_driven2D_nx707": dict(scenario="lid_driven2D", nx=707, mu=0.01, rho=1.0, dt=0.01),
    "lid_driven2D_nx1414": dict(scenario="lid_driven2D", nx=1414, mu=0.01, rho=1.0, dt=0.01),
    "lid_driven2D_nx2828": dict(scenario="lid_driven2D", nx=2828, mu=0.01, rho=1.0, dt=0.01),
    "lid_driven2D_nx64": dict(scenario="lid_driven2D", nx=64, mu=0.01, rho=1.0, dt=0.01),
    # BASELINE.json configs[2..4] on the mapped split-triangle stenosis channel
    "stenosis_backflow_1m": dict(scenario="stenosis_mesh_variable", res=0.03, dt=1e-3, v_max=100.0),
    "stenosis_backflow_4m": dict(scenario="stenosis_mesh_variable", res=0.015, dt=1e-3, v_max=100.0),
    "stenosis_pressure_4m": dict(scenario="stenosis_pressure", res=0.015, dt=1e-3, p_inlet=80.0, R_resistance=10.0),
    "stenosis_pressure_structured_16m": dict(scenario="stenosis_pressure_structured", res=0.0075, dt=1e-3,
                                             p_inlet=80.0, R_resistance=10.0),
    # the same transfinite grid with the reference's recombined Q1 cells
    "stenosis_pressure_structured_q1_1m": dict(scenario="stenosis_pressure_structured", res=0.0212, dt=1e-3,
                                               p_inlet=80.0, R_resistance=10.0, cell_type="quadrilateral"),
    "stenosis_pressure_structured_q1_4m": dict(scenario="stenosis_pressure_structured", res=0.0106, dt=1e-3,
                                               p_inlet=80.0, R_resistance=10.0, cell_type="quadrilateral"),
    "stenosis_pressure_structured_q1_8m": dict(scenario="stenosis_pressure_structured", res=0.0075, dt=1e-3,
                                               p_inlet=80.0, R_resistance=10.0, cell_type="quadrilateral"),
}
CPU_SAMPLE_NX = 128  # bounded CPU sample of the same workload (same physics, coarser mesh)

PROF_CLASSES = {0: "spmv_node(J)", 1: "cell_jacobian", 2: "gather_matrix", 3: "cell_residual",
                4: "cheb_step<2>(A00,l0)", 5: "cheb_step<1>(Lp,l0)", 6: "mdot", 7: "maxpy_norm",
                8: "numeric_rap<2>(l0)"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 8:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_scenario(w, **solver_kw):
    if w["scenario"] == "lid_driven2D":
        from cfd_hemodynamic_b200.src.scenarios.lid_driven2D import LidDriven2DSimulation
        return LidDriven2DSimulation("stabilized_schur", w["dt"], 1.0, rho=w["rho"], mu=w["mu"], nx=w["nx"], **solver_kw)
    if w["scenario"] == "stenosis_mesh_variable":
        from cfd_hemodynamic_b200.src.scenarios.stenosis_mesh_variable import StenosisMeshVariableSimulation
        return StenosisMeshVariableSimulation("stabilized_schur_backflow", w["dt"], 1.0, grade="severe",
                                              v_max=w["v_max"], res=w["res"], **solver_kw)
    from cfd_hemodynamic_b200.src.scenarios.stenosis_pressure_structured import StenosisPressureStructuredSimulation
    return StenosisPressureStructuredSimulation("stabilized_schur_pressure_backflow", w["dt"], 1.0, grade="severe",
                                                p_inlet=w["p_inlet"], R_resistance=w["R_resistance"], res=w["res"],
                                                cell_type=w.get("cell_type", "triangle"), **solver_kw)


def oracle_steps(w, nx, steps):
    """The CPU port (oracle/ns_oracle.py): same scenario, Newton + sparse LU, `steps`
    time steps at mesh size nx.  Returns (ndof, seconds)."""
    from cfd_hemodynamic_b200.fem import mesh as M
    from oracle import ns_oracle as O
    from tests import common as T
    mesh = M.create_unit_square(None, nx, nx)
    prob = T.make_problem(mesh, dt=w["dt"], rho=w["rho"], mu=w["mu"], f=(0.0, 0.0))
    x = prob.x
    n = prob.n
    ext = M.exterior_facet_indices(mesh.topology)
    prob.facet_sets = [O.FacetSet(pairs=mesh.topology.facet_cell_pairs(ext), a_p=1.0, a_g=1.0)]
    walls = np.nonzero(np.isclose(x[:, 0], 0) | np.isclose(x[:, 0], 1) | np.isclose(x[:, 1], 0))[0]
    lidf = M.locate_entities_boundary(mesh, 1, lambda X: np.isclose(X[1], 1.0) & (X[0] > 1e-10) & (X[0] < 1 - 1e-10))
    lid = np.unique(mesh.topology.facet_vertices[lidf])
    g1 = np.zeros(2 * n)
    g1[0::2] = 1.0
    prob.bcs = T.oracle_bcs(prob, [("u", walls, np.zeros(2 * n)), ("u", lid, g1)])
    from oracle.c_oracle import FastAssembler
    asm = FastAssembler(prob)            # C cell kernels, OpenMP over all host cores
    xk = np.zeros(3 * n)
    un = np.zeros(2 * n)
    t0 = time.perf_counter()
    for _ in range(steps):
        xk = O.remove_nullspace(prob, xk)
        xk, its, reason = O.newton_solve(prob, xk, un, asm=asm)
        un = xk[:2 * n].copy()
    return 3 * n, time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    if w["scenario"] != "lid_driven2D":
        w = WORKLOADS["lid_driven2D_nx707"]      # the CPU arm is only wired for the default workload
    cores = os.cpu_count() or 1
    for _ in range(min(args.warmup, 1)):
        oracle_steps(w, 16, 1)
    steps = max(1, args.steps)
    # bounded sample: the mesh is sized so that K steps finish within a few minutes
    # (sparse-LU Newton: ~30 s per step at nx=128, ~4 s at 64, ~1.5 s at 40, < 1 s at 32)
    sample_nx = 128 if steps <= 4 else 64 if steps <= 20 else 40 if steps <= 80 else 32
    ndof, secs = oracle_steps(w, sample_nx, steps)
    val = ndof * steps / secs
    line = {
        "impl": "reference", "metric": "DOF-timesteps/s", "value": val, "unit": "DOF-timesteps/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "solver": "stabilized_schur",
                   "note": "DOLFINx/PETSc are not installable in this image; CPU arm = oracle port "
                           "(C/OpenMP element kernels + SciPy SuperLU Newton) on a bounded sample of the workload"},
        "cpu_baseline": {"value": val, "unit": "DOF-timesteps/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} time step(s) of {w['scenario']} at nx={sample_nx} "
                                   f"({ndof} DOFs); assembly on all cores (OpenMP), sparse LU single-threaded"},
        "e2e": {"value": val, "unit": "DOF-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_distributed(args, w, W, K, world, rank, local_rank):
    """N > 1: ONE mesh partitioned over the GPUs (vertex-owned x-slabs + overlap, halo exchange
    and Krylov allreduces over NCCL).  Weak scaling: the per-GPU cell count is held at the
    single-GPU workload's, so the global mesh is nx*sqrt(N) squared."""
    import contextlib
    import torch
    import torch.distributed as dist
    from cfd_hemodynamic_b200.distributed_solver import DistributedStabilizedSchur
    from cfd_hemodynamic_b200.parallel import slab_partition
    from cfd_hemodynamic_b200.src.scenarios.lid_driven2D import LidDriven2DSimulation
    nx = int(round(w["nx"] * math.sqrt(world)))
    with contextlib.redirect_stdout(sys.stderr):
        sc = LidDriven2DSimulation("stabilized_schur", w["dt"], 1.0, rho=w["rho"], mu=w["mu"], nx=nx, host_only=True)
    tables = sc.solver.export_tables()
    owner = slab_partition(tables["x"][:, 0], world)
    ds = DistributedStabilizedSchur(tables, owner, local_rank)
    dev = ds.hemo.device
    ndof = 3 * ds.n_global
    E = tables["cells"].shape[0]

    def barrier():
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(W):
        ds.step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ds.hemo.launches + ds.hemo_p.launches
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    newton = ksp = 0
    e0.record()
    for _ in range(K):
        ds.step_device()
        newton += ds.its_snes
        ksp += ds.its_ksp
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    launches = torch.tensor([float(ds.hemo.launches + ds.hemo_p.launches - l0)], dtype=torch.float64, device=dev)
    dist.all_reduce(launches)
    clocks = sampler.stop()
    # e2e: owned values are pulled to pinned host memory and u_prev pushed back every step
    no, nl = ds.part.n_owned, ds.n
    host = torch.empty(3 * nl, dtype=torch.float64).pin_memory()
    hun = torch.empty(2 * nl, dtype=torch.float64).pin_memory()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        ds.d_un.copy_(hun.copy_(ds.d_un, non_blocking=True), non_blocking=True)   # host-owned time-level shift
        ds.step_device()
        host.copy_(ds.d_x, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    if rank == 0:
        peak, peak_kind = measured_peaks()
        line = {
            "metric": "DOF-timesteps/s", "value": ndof * K / (ms * 1e-3), "unit": "DOF-timesteps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "solver": "stabilized_schur", "cells": E, "dofs": ndof,
                       "global_nx": nx, "cells_per_gpu": E // world, "dt": w["dt"], "mu": w["mu"], "rho": w["rho"],
                       "newton_its_per_step": newton / K, "fgmres_its_per_step": ksp / K,
                       "parallelism": f"domain decomposition: {world} vertex-owned x-slabs, overlap {ds.overlap} cell "
                                      f"layers, all-gather halo ({ds.halo.bytes_per_update} B/update), "
                                      f"replicated global pressure V-cycle, NCCL allreduce for Krylov/Newton reductions",
                       "l2": "inputs larger than L2 (per-GPU matrix %.0f MB vs 126 MB L2)" % (8 * ds.hemo.nnz / 1e6)},
            "e2e": {"value": ndof * K / float(te.item()), "unit": "DOF-timesteps/s",
                    "h2d_bytes_per_step": world * 8 * 2 * nl, "d2h_bytes_per_step": world * 8 * 5 * nl},
            "gpu_launches": int(launches.item()), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "see the N=1 line (same kernels per partition)", "achieved": None,
                         "peak": peak, "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_kind},
            "cpu_baseline": None,
        }
        print(json.dumps(line))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lid_driven2D_nx707", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path exists only as sm_100a kernels (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W = max(3, args.warmup)
    K = max(1, args.steps)
    w = WORKLOADS[args.workload]

    import contextlib
    if world > 1:
        run_distributed(args, w, W, K, world, rank, local_rank)
        return
    with contextlib.redirect_stdout(sys.stderr):       # stdout carries exactly one JSON line
        sc = build_scenario(w, device=local_rank)
    s = sc.solver
    hemo = s.hemo
    ndof = s.N
    E = s._cells_host.shape[0]
    nnz = hemo.nnz
    dev = hemo.device

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- device-resident timed region --------------------------------------
    for _ in range(W):
        s.step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    hemo.prof_enable(True)
    launches0 = hemo.launches
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    newton = ksp = 0
    e0.record()
    for _ in range(K):
        s.step_device()
        newton += s.its_snes
        ksp += s.its_ksp
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = hemo.launches - launches0
    prof = {c: hemo.prof_get(c) for c in PROF_CLASSES}
    hemo.prof_enable(False)
    clocks = sampler.stop()
    # The preconditioner replays as a CUDA graph, so its kernels carry no event pairs in the
    # timed region above.  Two extra steps with direct launches (same kernels, same data) time
    # them with the same CUDA-event mechanism; their share is scaled to the timed region.
    hemo.use_graph(False)
    s.step_device()
    hemo.prof_enable(True)
    g0 = torch.cuda.Event(enable_timing=True)
    g1 = torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(2):
        s.step_device()
    g1.record()
    torch.cuda.synchronize(dev)
    ms_nograph = g0.elapsed_time(g1)
    for c in (4, 5):
        t, k = hemo.prof_get(c)
        prof[c] = (t * (ms / ms_nograph) if ms_nograph > 0 else t, int(round(k * K / 2)))
    hemo.prof_enable(False)
    hemo.use_graph(True)
    s.step_device()          # re-captures the graph
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    value = world * ndof * K / (ms * 1e-3)

    # ---- isolated assembly timings (cells/s, nnz/s) ------------------------------
    def time_fn(fn, reps=5):
        fn()
        torch.cuda.synchronize(dev)
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / reps

    jac_ms = time_fn(lambda: hemo.assemble_jacobian(s.d_x, s.d_un, s.d_vals))
    res_ms = time_fn(lambda: hemo.assemble_residual(s.d_x, s.d_un, s.d_bcval, s.d_g))

    # ---- end to end through the plugin API (host buffers) ------------------------
    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            s.solveStep()
            s.u_prev.x.array[:] = s.u_sol.x.array[:]
            s.p_prev.x.array[:] = s.p_sol.x.array[:]
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            s.solveStep()
            s.u_prev.x.array[:] = s.u_sol.x.array[:]
            s.p_prev.x.array[:] = s.p_sol.x.array[:]
        barrier()
        t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e = {"value": world * ndof * K / float(t_e2e.item()), "unit": "DOF-timesteps/s",
               "h2d_bytes_per_step": s.h2d_bytes_per_step, "d2h_bytes_per_step": s.d2h_bytes_per_step}

    # ---- roofline of the dominant kernel -------------------------------------------
    n = s.n
    nnz_node = hemo.nnz_node
    nv = int(s._cells_host.shape[1])                      # 3: P1 triangles, 4: Q1 quadrilaterals
    ae_bytes = 72 * nv * nv * E                           # element matrices, 9 doubles per node pair
    alg_bytes = {
        0: 76 * nnz_node + 52 * n,                        # J values + node cols + rowptr + x + y
        1: 4 * nv * E + 8 * E + 56 * n + ae_bytes,        # cells, h, nodal gathers, element matrices out
        2: ae_bytes + 4 * nv * nv * E + 12 * nnz_node + 72 * nnz_node,
        3: 4 * nv * E + 8 * E + 56 * n + 24 * nv * E,
        4: 20 * nnz_node + 4 * n + 56 * n,                # A00 BSR2 fp32 values + cols + rowptr + 7 fp32 vectors of 2n
        5: 8 * nnz_node + 4 * n + 28 * n,
        6: None, 7: None,
        8: None,
    }
    peak, peak_kind = measured_peaks()
    # the cell kernels are FP64-pipe bound at the reference quadrature (on Q1 cells by two
    # orders of magnitude): the HBM roofline line is chosen among the bandwidth-bound kernel classes
    hbm_classes = (0, 2, 4, 5)
    dom = max((c for c in hbm_classes if prof[c][1] > 0 and alg_bytes[c]), key=lambda c: prof[c][0], default=0)
    dms, dcnt = prof[dom]
    achieved = alg_bytes[dom] / (dms / dcnt * 1e-3) / 1e9 if dcnt else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("workload") == args.workload:
            traffic = tj.get("dram_bytes_per_launch", {}).get(PROF_CLASSES[dom])
    roofline = {"bound": "hbm", "kernel": PROF_CLASSES[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_kind,
                "launches_timed": dcnt, "avg_us": 1e3 * dms / max(dcnt, 1),
                "share_of_step": dms / ms,
                "fp64_bound": {"kernels": ["cell_jacobian", "cell_residual"],
                               "evidence": "ncu sm__pipe_fp64_cycles_active: 75 % (Q1 J_uu work items, "
                                           "profiles/r01_q1_ncu_full_cell_kernels.csv); DRAM <= 10 %"},
                "other_kernels": {PROF_CLASSES[c]: {"ms_total": prof[c][0], "launches": prof[c][1],
                                                    "GBps": (alg_bytes[c] / (prof[c][0] / prof[c][1] * 1e-3) / 1e9
                                                             if alg_bytes[c] and prof[c][1] else None)}
                                  for c in prof if c != dom}}

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on a bounded sample --------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and w["scenario"] == "lid_driven2D":
        cdof, csecs = oracle_steps(w, CPU_SAMPLE_NX, 1)
        cpu = {"value": cdof / csecs, "unit": "DOF-timesteps/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"1 time step of {w['scenario']} at nx={CPU_SAMPLE_NX} ({cdof} DOFs) with the numpy/SciPy "
                         f"oracle port: C/OpenMP cell kernels on all cores + SciPy SuperLU Newton ({csecs:.1f} s); "
                         f"DOLFINx/PETSc not installable here"}

    if rank == 0:
        line = {
            "metric": "DOF-timesteps/s", "value": value, "unit": "DOF-timesteps/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "solver": sc.solver_name, "cells": E, "dofs": ndof,
                       "nnz": nnz, "dt": w["dt"], "mu": float(s.mu.value), "rho": float(s.rho.value),
                       "newton_its_per_step": newton / K, "fgmres_its_per_step": ksp / K,
                       "parallelism": "1 GPU" if world == 1 else f"{world} independent replicas (one mesh per GPU)",
                       "l2": "inputs larger than L2 (matrix %.0f MB, element buffer %.0f MB vs 126 MB L2)"
                             % (8 * nnz / 1e6, 72 * int(s._cells_host.shape[1]) ** 2 * E / 1e6),
                       "cell_type": s.mesh.topology.cell_name()},
            "assembly": {"jacobian_cells_per_s": world * E / (jac_ms * 1e-3), "jacobian_nnz_per_s": world * nnz / (jac_ms * 1e-3),
                         "residual_cells_per_s": world * E / (res_ms * 1e-3), "jacobian_ms": jac_ms, "residual_ms": res_ms},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
