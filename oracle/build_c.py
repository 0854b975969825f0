"""Build the C restatement of the cell kernels (oracle/c/p1tri_cells.c) into oracle/_build/.
Run by __graft_entry__.build(); test/baseline infrastructure only."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "p1tri_cells.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libhemo_ref_cells.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) > os.path.getmtime(SRC):
        return OUT
    cmd = ["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", "-o", OUT, SRC, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("gcc failed building the C oracle")
    return OUT


if __name__ == "__main__":
    print(build(force=True))
