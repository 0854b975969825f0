"""Pressure-driven stenosis (reference src/scenarios/stenosis_pressure.py): same
boundary model as the structured variant; the reference meshes it with an
unstructured gmsh triangulation, here the mapped split-triangle generator is
used with `res` derived from `n_elements_radial`."""
from .stenosis_pressure_structured import StenosisPressureStructuredSimulation


class StenosisPressureSimulation(StenosisPressureStructuredSimulation):
    scenario_name = "stenosis_pressure"
    default_cell_type = "triangle"          # the reference's gmsh mesh here is not recombined

    def __init__(self, solver_name, dt, T, f=(0, 0), grade="severe", p_inlet: float = 80.0,
                 R_resistance: float = None, v_max: float = None, *, rho: float = 1.060e-3, mu: float = 3.5e-3,
                 n_elements_radial: int = None, **kwargs):
        if n_elements_radial is not None and "res" not in kwargs:
            kwargs["res"] = 2.0 * 1.57 / (2 * int(n_elements_radial))
        super().__init__(solver_name, dt, T, f, grade, p_inlet, R_resistance, v_max, rho=rho, mu=mu, **kwargs)
