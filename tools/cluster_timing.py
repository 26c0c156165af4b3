"""Phase timestamps (clock64) of one CTA of the cluster smoother on level 1: python tools/cluster_timing.py [cta]
(needs a library built with NVCC_EXTRA=-DFSB_DEBUG_STAMPS python sci-solver_fem_b200/build.py -f)"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sci_solver_fem_b200 as fsb
cta = int(sys.argv[1]) if len(sys.argv) > 1 else 0
v, t = fsb.meshio.kuhn_cube(118)
s = fsb.FEMSolver.from_arrays(v, t)
s.solverType_, s.tolerance_, s.maxIters_, s.seed_, s.useGraphs_ = 1, 1e-8, 3, 0, 0
s.setup()
b = np.random.default_rng(1).uniform(-1, 1, len(v))
L = s._L
L.fsb_debug_stamps.argtypes = [C.c_int, C.c_void_p]
s.solve(np.zeros_like(b), b)
L.fsb_debug_stamps(cta + 1, None)
s.maxIters_ = 1
s.solve(np.zeros_like(b), b)   # last cluster-kernel launch = post-smooth of the coarsest cluster level; all overwrite: last writer wins
out = np.zeros(64, dtype=np.int64)
L.fsb_debug_stamps(0, out.ctypes.data_as(C.c_void_p))
t0 = out[0]
names = {0: "start", 1: "init done", 2: "staged", 3: "first cluster.sync done", 40: "sweeps done", 41: "end"}
for k in range(64):
    if out[k]:
        lab = names.get(k, f"sweep {(k-4)//2} {'computed' if k % 2 == 0 else 'synced'}")
        print(f"{k:3d} {lab:28s} {out[k]-t0:10d} cycles")
