"""Per-kernel CUDA-event profile of one solve (no ncu): python tools/quick_profile.py [--cube N]"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sci_solver_fem_b200 as fsb
ap = argparse.ArgumentParser(); ap.add_argument("--cube", type=int, default=118); ap.add_argument("--inner", type=int, default=5); ap.add_argument("--maxiters", type=int, default=200); args = ap.parse_args()
v, t = fsb.meshio.kuhn_cube(args.cube)
s = fsb.FEMSolver.from_arrays(v, t)
s.solverType_, s.tolerance_, s.maxIters_, s.seed_ = 1, 1e-8, args.maxiters, 0
s.preInnerIters_ = s.postInnerIters_ = args.inner
s.setup()
b = np.random.default_rng(1234).uniform(-1, 1, len(v))
for _ in range(3):
    s.solve(np.zeros_like(b), b)
ms = s.time_ms("solve"); it = s.iterations
s.profile_ = 1
s.solve(np.zeros_like(b), b)
prof = s.profile_report()
tot = sum(m for _, m in prof.values())
print(f"cube {args.cube}: solve {ms:.2f} ms, {it} iterations, {ms/it*1e3:.0f} us/iter (graph); profiled sum {tot:.2f} ms; setup {s.time_ms('setup'):.1f} ms")
print("  setup phases (ms):", {k: round(s.time_ms(k), 2) for k in ("pattern", "assemble", "setup", "setup_aggregation", "setup_permute_split", "setup_prolongator", "setup_galerkin", "setup_galerkin_AP_L0", "setup_galerkin_RAP_L0", "setup_galerkin_AP_L1", "setup_galerkin_RAP_L1", "setup_galerkin_RAP_L2", "setup_coarse_inverse")})
for (k, l), (c, m) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"  {k:14s} L{l}  {c:4d} launches  {m/c*1e3:8.1f} us each  {100*m/tot:5.1f}%")
