"""Minimal driver for ncu: builds the N-cube problem and runs `--solves` PCG solves (no timing here).

    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'smooth|csr_|spmv|cg_|dot_|coarse' \
        --csv --log-file gpurun_out/launches.csv python tools/profile_solve.py --cube 118
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sci_solver_fem_b200 as fsb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cube", type=int, default=118)
ap.add_argument("--solves", type=int, default=1)
ap.add_argument("--graphs", type=int, default=0)
ap.add_argument("--maxiters", type=int, default=200)
args = ap.parse_args()
v, t = fsb.meshio.kuhn_cube(args.cube)
s = fsb.FEMSolver.from_arrays(v, t)
s.solverType_, s.tolerance_, s.maxIters_, s.seed_, s.useGraphs_ = 1, 1e-8, args.maxiters, 0, args.graphs
s.setup()
rng = np.random.default_rng(1234)
b = rng.uniform(-1, 1, len(v))
for _ in range(args.solves):
    x = s.solve(np.zeros_like(b), b)
print("iterations", s.iterations, "relres", s.relres, "levels", [(s.level_rows(l), s.level_nnz(l)) for l in range(s.num_levels())],
      "solve_ms", s.time_ms("solve"), "setup_ms", s.time_ms("setup"))
