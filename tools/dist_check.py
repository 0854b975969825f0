"""Partition invariance on real GPUs (run under torchrun, one rank per GPU):
the N-partition solve must reproduce the single-GPU solve of the same mesh.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import contextlib
import numpy as np, torch, torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from cfd_hemodynamic_b200.src.scenarios.lid_driven2D import LidDriven2DSimulation
from cfd_hemodynamic_b200.distributed_solver import DistributedStabilizedSchur
from cfd_hemodynamic_b200.parallel import slab_partition

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 24
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
case = sys.argv[3] if len(sys.argv) > 3 else "lid"
cell_type = sys.argv[4] if len(sys.argv) > 4 else "triangle"
tight = dict(snes_rtol=1e-11, snes_stol=0.0, ksp_rtol=1e-9, ksp_restart=100) if nx <= 64 else {}
def make(host_only, **kw):
    if case == "lid":
        return LidDriven2DSimulation("stabilized_schur", 0.01, 1.0, rho=1, mu=0.01, nx=nx, cell_type=cell_type,
                                     host_only=host_only, **kw)
    if case == "pressure":
        from cfd_hemodynamic_b200.src.scenarios.stenosis_pressure_structured import StenosisPressureStructuredSimulation
        return StenosisPressureStructuredSimulation("stabilized_schur_pressure_backflow", 0.005, 1.0, grade="moderate", cell_type=cell_type,
                                                    p_inlet=2.0, R_resistance=50.0, res=3.14 / nx, L=20.0,
                                                    x_position_stenosis=8.0, schur_mode="laplace",
                                                    host_only=host_only, **kw)
    from cfd_hemodynamic_b200.src.scenarios.stenosis_mesh_variable import StenosisMeshVariableSimulation
    # stenosis channel with Dirichlet inlet + backflow-stabilised open outlet (nx = cells across the inlet)
    return StenosisMeshVariableSimulation("stabilized_schur_backflow", 0.005, 1.0, grade="moderate", v_max=5.0,
                                          res=3.14 / nx, L=20.0, x_position_stenosis=8.0, schur_mode="laplace",
                                          host_only=host_only, **kw)

with contextlib.redirect_stdout(sys.stderr):
    sc = make(True, **tight)
tables = sc.solver.export_tables()
owner = slab_partition(tables["x"][:, 0], world)
ds = DistributedStabilizedSchur(tables, owner, lr, verbose=bool(os.environ.get('DIST_VERBOSE')),
                                overlap=int(os.environ['DIST_OVERLAP']) if os.environ.get('DIST_OVERLAP') else None,
                                coarse_pressure=not os.environ.get('DIST_NO_COARSE'),
                                coarse_level=int(os.environ.get('DIST_COARSE_LEVEL', '2')))
torch.cuda.synchronize(); dist.barrier()
t0 = time.time()
its = []
for k in range(steps):
    ds.step_device()
    its.append((ds.its_snes, ds.its_ksp))
torch.cuda.synchronize(); dist.barrier()
dt = (time.time() - t0) / steps
u, p = ds.gather_solution()
if rank == 0:
    print(f"distributed: world {world} nx {nx} ms/step {1e3*dt:.1f} its {its} {ds.comm_summary()}")
    if nx <= 128:
        with contextlib.redirect_stdout(sys.stderr):
            ref = make(False, device=lr, **tight)
        s = ref.solver
        for k in range(steps):
            s.step_device()
        if case == "pressure":
            print(f"outlet pressure p_c: serial {s._p_c:.10e} distributed {ds._outlet['p_c']:.10e}")
        x = s.d_x.cpu().numpy(); n = s.n
        ur, pr = x[:2 * n], x[2 * n:]
        eu = np.linalg.norm(u - ur) / np.linalg.norm(ur)
        ep = np.linalg.norm((p - p.mean()) - (pr - pr.mean())) / max(np.linalg.norm(pr - pr.mean()), 1e-300)
        print(f"partition invariance: rel err u {eu:.3e} p {ep:.3e} (serial its {s.its_snes},{s.its_ksp})")
        assert eu < 1e-8 and ep < 1e-8, (eu, ep)
dist.barrier()
dist.destroy_process_group()
