"""Solve time with and without the per-iteration CUDA graph: python tools/graph_vs_stream.py [--cube N]"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sci_solver_fem_b200 as fsb
ap = argparse.ArgumentParser(); ap.add_argument("--cube", type=int, default=118); args = ap.parse_args()
v, t = fsb.meshio.kuhn_cube(args.cube)
s = fsb.FEMSolver.from_arrays(v, t)
s.solverType_, s.tolerance_, s.maxIters_, s.seed_ = 1, 1e-8, 200, 0
s.setup()
b = np.random.default_rng(1234).uniform(-1, 1, len(v))
for g in (1, 0, 1, 0):
    s.useGraphs_ = g
    ts = []
    for _ in range(4):
        s.solve(np.zeros_like(b), b); ts.append(s.time_ms("solve"))
    print(f"graphs={g}: solve {min(ts):.2f} ms ({s.iterations} iterations, {min(ts)/s.iterations*1e3:.0f} us/iter)  all={['%.2f' % x for x in ts]}")
