"""sci-solver_fem_b200 — B200-native rebuild of SCI-Solver_FEM's solve path.

The directory name carries a hyphen (fixed by the project layout), so import it through the
root-level shim:  `import sci_solver_fem_b200 as fsb`.

Contents: `femsolver` (FEMSolver mirror over the C-ABI of libfemsolver_b200.so), `meshio`
(mesh / .mat formats either side of the path and the synthetic Kuhn-cube generator),
`build` (in-tree nvcc build for sm_100a).  The CUDA library is the only implementation of the
path: there is no CPU fallback and nothing here imports oracle/.
"""
from . import meshio  # noqa: F401
from .femsolver import (EXPORTED_SYMBOLS, FEMSolver, FEMSolverError, exchange_handles_torch, load_library,  # noqa: F401
                        split_by_weight)
