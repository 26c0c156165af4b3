"""Phase timestamps (clock64) of one CTA of the fine-level ELL smoother: python tools/ell_timing.py [cta] [kind]
kind 1: fine-level register-resident kernel (256-thread class), kind 3: shared-memory kernel of the coarser levels.
Needs a library built with NVCC_EXTRA=-DFSB_DEBUG_STAMPS python sci-solver_fem_b200/build.py -f"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sci_solver_fem_b200 as fsb
cta = int(sys.argv[1]) if len(sys.argv) > 1 else 0
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 1
v, t = fsb.meshio.kuhn_cube(118)
s = fsb.FEMSolver.from_arrays(v, t)
s.solverType_, s.tolerance_, s.maxIters_, s.seed_, s.useGraphs_ = 1, 1e-8, 3, 0, 0
s.setup()
b = np.random.default_rng(1).uniform(-1, 1, len(v))
L = s._L
L.fsb_debug_stamps.argtypes = [C.c_int, C.c_void_p]
s.solve(np.zeros_like(b), b)
L.fsb_debug_stamps((cta + 1) | (kind << 24), None)
s.maxIters_ = 1
s.solve(np.zeros_like(b), b)   # last stamped launch wins: the post-smoothing stage (kind 3: of the first level that uses that kernel ... last one launched)
out = np.zeros(64, dtype=np.int64)
L.fsb_debug_stamps(0, out.ctypes.data_as(C.c_void_p))
nz = [k for k in range(64) if out[k]]
t0 = out[nz[0]]
prev = t0
for k in nz:
    print(f"{k:3d} {out[k]-t0:10d} cycles  (+{out[k]-prev})")
    prev = out[k]
