"""Regenerates tests/golden/oracle_cube118_pcg.json: the CPU oracle's PCG solve of BASELINE configs[2]
(Kuhn cube N=118, b = A * egg-carton, defaults + solverType_=1, tolerance_=1e-8, seed_=0) — iteration count,
residual history, solution norm and samples.  About 30 s on 8 cores; the `-m gpu` test
tests/test_gpu_parity.py::test_config3_matches_the_oracle_golden compares the CUDA path against it without
having to run the oracle at that size on the GPU box.

    python tests/golden/make_oracle_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import oracle as orc  # noqa: E402

N = 118
verts, tets, xstar = bench.build_problem(N)
o = orc.Oracle(64, solverType=1, tolerance=1e-8, maxIters=200, seed=0)
o.pattern(len(verts), tets); o.assemble(verts); o.setup()
b = o.spmv(xstar)
x, it = o.solve(b)
h = np.array(o.resid_history())
idx = [0, 1000, 123456, 842579, 1685158]
g = {"cube": N, "iterations": int(it), "relres": float(h[-1]), "resid_history": [float(v) for v in h],
     "x_norm2": float(np.linalg.norm(x)), "b_norm2": float(np.linalg.norm(b)), "sample_idx": idx,
     "x_samples": [float(x[i]) for i in idx], "err_vs_exact": float(np.linalg.norm(x - xstar) / np.linalg.norm(xstar)),
     "levels": [int(o.level_rows(l)) for l in range(4)]}
json.dump(g, open(os.path.join(ROOT, "tests", "golden", "oracle_cube118_pcg.json"), "w"), indent=1)
print(g["iterations"], g["relres"], g["x_norm2"], g["err_vs_exact"], g["levels"])
