"""One cold pass of pattern + assembly + AMG setup (for an ncu launch list of the setup-phase kernels):
   ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file out.csv python tools/setup_only.py [--cube 118]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sci_solver_fem_b200 as fsb
ap = argparse.ArgumentParser(); ap.add_argument("--cube", type=int, default=118); ap.add_argument("--repeat", type=int, default=1); a = ap.parse_args()
v, t = fsb.meshio.kuhn_cube(a.cube)
s = fsb.FEMSolver.from_arrays(v, t)
s.solverType_, s.seed_ = 1, 0
for r in range(a.repeat - 1):   # warm-pool stage timings of the rebuilds
    s.getMatrixFromMesh(); print("rebuild", {k: round(s.time_ms(k), 2) for k in ("pattern", "assemble")}, flush=True)
for r in range(a.repeat):   # --repeat 2 with FSB_SETUP_TRACE=1: the second pass shows the warm-pool laps
    print(f"-- pass {r}", file=sys.stderr, flush=True)
    s.setup()
keys = ("pattern", "assemble", "setup", "setup_aggregation", "setup_permute_split", "setup_prolongator", "setup_galerkin",
        "setup_coarse_inverse", "setup_dense_tail", "setup_block_smoothers")
print({k: round(s.time_ms(k), 2) for k in keys})
