"""bench.py contract on the CPU side: `--impl reference` prints exactly ONE JSON line on stdout with the driver's keys, the same
`config` function serves both arms, and non-zero ranks of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    e["OMP_NUM_THREADS"] = "2"
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "lid_driven2D_nx64",
                           "--steps", "1", "--warmup", "0", *extra], capture_output=True, text=True, env=e, timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run([])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert KEYS <= set(line)
    assert line["impl"] == "reference" and line["metric"] == "DOF-timesteps/s" and line["higher_is_better"] is True
    assert line["config"]["workload"] == "lid_driven2D_nx64" and line["config"]["dofs"] == 3 * 65 * 65
    assert line["steps"] == 1 and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    r = _run(["--gpus", "2"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_follows_the_weak_scaled_mesh_of_an_n_gpu_run():
    """At N > 1 the GPU arm solves one cavity of nx * sqrt(N) squared: rank 0 of the reference arm marches the same mesh."""
    r = _run(["--gpus", "4"], env={"RANK": "0", "LOCAL_RANK": "0", "WORLD_SIZE": "4"})
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.strip()][-1])
    assert line["n_gpus"] == 4 and line["config"]["global_nx"] == 128 and line["config"]["dofs"] == 3 * 129 * 129
    assert line["config"]["cells_per_gpu"] == 2 * 128 * 128 // 4
