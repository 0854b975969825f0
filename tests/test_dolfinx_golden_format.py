"""The reader of tests/test_dolfinx_golden.py on a synthetic file in the dump format (no DOLFINx needed)."""
import numpy as np

from tests import test_dolfinx_golden as G


def test_reader_on_synthetic_file(tmp_path):
    path = str(tmp_path / "dolfinx_synthetic.npz")
    G._write_synthetic(path)
    G.test_oracle_matches_dolfinx(path)
    g = np.load(path)
    prob = G.problem_from_golden(g)
    assert prob.n == 30 and len(prob.bcs) == 2
