import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np, sci_solver_fem_b200 as fsb
v,t = fsb.meshio.kuhn_cube(118)
s = fsb.FEMSolver.from_arrays(v,t)
s.solverType_, s.tolerance_, s.maxIters_, s.seed_ = 1, 1e-8, 200, 0
b = np.random.default_rng(1).uniform(-1,1,len(v)); x=np.zeros_like(b)
for i in range(5):
    t0=time.perf_counter(); s.solveFEM(x.copy(), b); dt=time.perf_counter()-t0
    print('solveFEM wall %.1f ms'%(dt*1e3), 'setup', round(s.time_ms('setup'),1), 'solve', round(s.time_ms('solve'),1), {k: round(s.time_ms(k),1) for k in ('setup_dense_tail','setup_galerkin','setup_aggregation','setup_permute_split','setup_RA_L1')})
