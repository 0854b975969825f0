"""Build the C restatements under oracle/c/ into oracle/_build/ (test / baseline infrastructure only):
p1tri_cells.c (FFCx-style cell kernels) and ref_ksp.c (the reference's PETSc solver configuration on
host cores).  Run by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
TARGETS = {
    "cells": (os.path.join(HERE, "c", "p1tri_cells.c"), os.path.join(OUT_DIR, "libhemo_ref_cells.so")),
    "ksp": (os.path.join(HERE, "c", "ref_ksp.c"), os.path.join(OUT_DIR, "libhemo_ref_ksp.so")),
}
SRC, OUT = TARGETS["cells"]


def _cpu_has_avx2():
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return False
    return " avx2" in flags and " fma" in flags


def _arch_flags():
    # x86-64-v3 (AVX2 + FMA): what a DOLFINx user gets from the recommended "-O3 -march=native" JIT flags on
    # any current server CPU, without tying the prebuilt library to this container's exact model
    return ["-march=x86-64-v3"] if _cpu_has_avx2() else []


def build(force=False, which="cells"):
    src, out = TARGETS[which]
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = out + ".flags"
    want = " ".join(_arch_flags())
    have = open(stamp).read() if os.path.exists(stamp) else None
    if not force and os.path.exists(out) and os.path.getmtime(out) > os.path.getmtime(src) and have == want:
        return out
    cmd = ["gcc", "-O3", *_arch_flags(), "-fno-math-errno", "-fopenmp", "-fPIC", "-shared", "-o", out, src, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("gcc failed building " + src)
    with open(stamp, "w") as fh:
        fh.write(want)
    return out


if __name__ == "__main__":
    for k in TARGETS:
        print(build(force=True, which=k))
