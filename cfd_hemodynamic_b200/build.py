"""Build libhemo_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).
Every translation unit is compiled to an object in csrc/_obj/ (in parallel, rebuilt only when it or a header is
newer), then linked."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libhemo_sm100.so")
SOURCES = ["assembly.cu", "assembly_q1.cu", "assembly_p2.cu", "assembly_tet.cu", "assembly_curlcurl.cu", "postproc.cu", "linalg.cu", "amg.cu", "solver.cu",
           "krylov.cu", "comm.cu", "host_setup.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--extended-lambda",
         "-Xcompiler", "-fPIC"]


def _headers_mtime() -> float:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "hemo.h"))
    return max(os.path.getmtime(h) for h in hs)


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.startswith("_")]
    deps.append(os.path.join(HERE, "..", "include", "hemo.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hm = _headers_mtime()

    def compile_one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if not force and os.path.exists(o) and os.path.getmtime(o) > max(os.path.getmtime(s), hm):
            return o, 0, ""
        cmd = [nvcc, *FLAGS, "-c", s, "-o", o] + (["-Xptxas", "-v"] if verbose else [])
        r = subprocess.run(cmd, capture_output=True, text=True)
        return o, r.returncode, r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    bad = [r for r in results if r[1] != 0]
    for o, rc, out in results:
        if out and (verbose or rc != 0):
            sys.stderr.write(out)
    if bad:
        raise RuntimeError("nvcc failed building libhemo_sm100.so")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + [r[0] for r in results] + ["-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libhemo_sm100.so")
    return LIB


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
