import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np, sci_solver_fem_b200 as fsb
v,t = fsb.meshio.kuhn_cube(118)
t0=time.time(); s = fsb.FEMSolver.from_arrays(v,t); print('construct wall', time.time()-t0, 'pattern', s.time_ms('pattern'), 'assemble', s.time_ms('assemble'))
for i in range(3):
    t0=time.time(); s.getMatrixFromMesh(); print('re-assemble wall', time.time()-t0, 'pattern', s.time_ms('pattern'), 'assemble', s.time_ms('assemble'))
for i in range(3):
    t0=time.time(); s.setup(); print('setup wall', time.time()-t0, 'setup', s.time_ms('setup'))
