"""Partition invariance on real GPUs (this check lived in tools/ only): the 2-partition solve — NCCL halo
exchange and allreduces inside the library, distributed FGMRES, two-level Schwarz preconditioner — must reproduce the
single-GPU solution of the same mesh to 1e-8 (tight solver tolerances on both sides).  Skipped with fewer than 2 GPUs."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(n, *args):
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "dist_check.py"), *map(str, args)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "partition invariance" in r.stdout, r.stdout[-2000:]
    return r.stdout


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("case", ["lid", "stenosis", "pressure"])
def test_two_partitions_reproduce_single_gpu(case):
    out = _run(2, 32, 3, case)
    assert "nccl_version" in out
