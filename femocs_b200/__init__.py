"""femocs_b200 -- B200 (sm_100a) implementation of FEMOCS' per-step electrostatic hot path.

Product code lives in csrc/ (CUDA kernels + C ABI, built into lib/libfemocs_b200.so) and in
solver.py (host-side mirror of the reference's PoissonSolver / Interpolator / SolutionReader /
Pic interface).  There is no CPU fallback."""
from .solver import (Context, CurrentHeatSolver, FemocsB200Error, HeatingConfig, FieldConfig, FieldReader, Interpolator, PartitionPlan, Pic, PoissonSolver,  # noqa: F401
                     SolutionReader, PRECOND_CHEBYSHEV, PRECOND_JACOBI, PRECOND_TWOLEVEL)
