"""`--solver stabilized_schur` on B200: SUPG/PSPG/LSIC-stabilised P1–P1
Navier–Stokes, Newton, FGMRES with a Schur-complement block preconditioner.
Triangles, quadrilaterals and (through `_stabilized_tet`) tetrahedra.

Drop-in for reference src/solvers/stabilized_schur.py: same module-level
`Solver` class, same constructor / `setup` / `solveStep` contract, loaded by
name from `Scenario.__init__` (reference src/scenario.py:63-78).
"""
from typing import Callable

import numpy as np

from ._stabilized_tet import StabilizedSchurTetB200


class Solver(StabilizedSchurTetB200):
    MAX_ITER = 20
    variant = "schur"

    def __init__(self, mesh, dt: float, rho: float, mu: float, f: list,
                 initial_velocity: Callable[[np.ndarray], np.ndarray] = None, **kwargs):
        # reference signature: stabilized_schur.py:43-52 (unknown kwargs are accepted and ignored)
        super().__init__(mesh, dt, rho, mu, f, initial_velocity, **kwargs)
