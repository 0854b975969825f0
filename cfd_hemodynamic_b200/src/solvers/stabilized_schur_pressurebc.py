"""`--solver stabilized_schur_pressurebc` on B200 (reference src/solvers/stabilized_schur_pressurebc.py): the
curl-curl / rotational formulation — viscous term mu (curl u_m . curl v), convection rho (curl u_m x u_m) . v -
rho |u_m|^2 / 2 div v, SUPG/PSPG/LSIC with the viscous part of the strong residual dropped (:85-160) — with natural
pressure conditions on inlet and outlet, `p (n.v)` with HALF the value passed (:63-64,189-190), and the Nitsche terms
for u_T = 0 written with curl x n (:193-201).  Wall velocities are the only Dirichlet conditions; `bcp` is ignored
(:219-220).

`setup()` appends the boundary terms to F each time it is called (`self.F += ...`, :189-201); the multiplicity is
reproduced like in the other variants.  The reference's sub-solves are `lu` (:270-274); here the
same block preconditioner as the rest of the family — parity is on the converged solution.

The library assembles this form with hemo_set_formulation(HEMO_FORM_CURLCURL) (csrc/assembly_curlcurl.cu) on P1
triangles and tetrahedra; everything after the element tensors is shared with stabilized_schur.
"""
from typing import Callable

import numpy as np

from ._stabilized_common import SET_INLET, SET_OUTLET
from ._stabilized_tet import StabilizedSchurTetB200


class Solver(StabilizedSchurTetB200):
    MAX_ITER = 20
    variant = "pressurebc"
    formulation = "curlcurl"
    _supported_cells = ("triangle", "tetrahedron")

    def __init__(self, mesh, dt: float, rho: float, mu: float, f: list,
                 initial_velocity: Callable[[np.ndarray], np.ndarray] = None,
                 p_inlet: float = None, p_outlet: float = None, beta_nitsche: float = 100.0, p_grade: int = 1, **kwargs):
        if p_inlet is None or p_outlet is None:
            raise ValueError("p_inlet and p_outlet are required for stabilized_schur_pressurebc. "
                             "Pass them via CLI: --p_inlet <value> --p_outlet <value>")
        if int(p_grade) != 1:
            raise NotImplementedError("stabilized_schur_pressurebc: the curl-curl kernels are written for P1-P1")
        self._p_inlet_val = float(p_inlet) / 2          # (:63-64)
        self._p_outlet_val = float(p_outlet) / 2
        self.beta_nitsche = float(beta_nitsche)
        super().__init__(mesh, dt, rho, mu, f, initial_velocity, p_grade=1, **kwargs)

    def _facet_setup(self, facet_tags, tags):
        c = float(self._setup_count)
        self._register_facets(SET_INLET, facet_tags.find(tags["inlet"]), pconst=c * self._p_inlet_val, a_n=c,
                              beta_n=self.beta_nitsche)
        self._register_facets(SET_OUTLET, facet_tags.find(tags["outlet"]), pconst=c * self._p_outlet_val, a_n=c,
                              beta_n=self.beta_nitsche)
