"""Host-side mirror of the reference's plugin API (reference src/): the same
module names (`solverBase`, `boundaryCondition`, `scenario`, `solvers.*`,
`scenarios.*`) so that scenario code written against the reference imports
from here unchanged, with DOLFINx objects replaced by the numpy shim in
`cfd_hemodynamic_b200.fem` when DOLFINx is absent."""
