"""Shared helpers for the parity tests (the oracle is imported here and ONLY by tests/bench/smoke)."""
import os

import numpy as np

import sci_solver_fem_b200 as fsb
from oracle.oracle import Oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def kuhn(N, **kw):
    return fsb.meshio.kuhn_cube(N, **kw)


def egg_carton(verts):
    return np.sin(2 * np.pi * verts[:, 0]) * np.sin(2 * np.pi * verts[:, 1]) * np.sin(2 * np.pi * verts[:, 2])


def make_oracle(verts, elems, labels=None, precision=64, **params):
    o = Oracle(precision, **params)
    ptr, col = o.pattern(len(verts), elems)
    val = o.assemble(verts, labels)
    return o, ptr, col, val


PARAM_MAP = dict(solverType="solverType_", tolerance="tolerance_", maxIters="maxIters_", seed="seed_", topSize="topSize_",
                 preInnerIters="preInnerIters_", postInnerIters="postInnerIters_", postRelaxes="postRelaxes_",
                 partitionMaxSize="partitionMaxSize_", randMisParameters="randMisParameters_", smootherWeight="smootherWeight_",
                 proOmega="proOmega_", maxLevels="maxLevels_", refLevel0NoPerm="refLevel0NoPerm_", aggregatorType="aggregatorType_")


def make_gpu(verts, elems, labels=None, **params):
    s = fsb.FEMSolver.from_arrays(verts, elems, labels)
    for k, v in params.items():
        setattr(s, PARAM_MAP[k], v)
    return s


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)
