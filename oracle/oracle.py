"""ctypes wrapper around oracle/libfem_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False):
    so = os.path.join(_HERE, "libfem_oracle.so")
    src = os.path.join(_HERE, "fem_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libfem_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_param.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.orc_pattern.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_get_pattern.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_assemble_tet.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_int]
        L.orc_assemble_tri.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 3 + [C.c_int]
        L.orc_get_values.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_set_values.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_tet_mass_integrals.argtypes = [C.c_void_p]
        L.orc_setup.argtypes = [C.c_void_p]
        L.orc_level_int.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p]
        L.orc_level_val.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p]
        L.orc_num_levels.argtypes = [C.c_void_p]
        L.orc_level_rows.argtypes = [C.c_void_p, C.c_int]
        L.orc_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_precond_permuted.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_final_relres.restype = C.c_double
        L.orc_final_relres.argtypes = [C.c_void_p]
        L.orc_resid_history.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_time.restype = C.c_double
        L.orc_time.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_spmv.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_randomized_mis.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_uint, C.c_void_p]
        L.orc_set_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """CPU restatement of the reference path.  precision=64: fp64 everywhere (north star);
    precision=32: the reference's mixed precision (fp32 hierarchy, fp64 PCG)."""

    PARAMS = ("maxLevels", "maxIters", "preInnerIters", "postInnerIters", "postRelaxes", "topSize",
              "randMisParameters", "partitionMaxSize", "aggregatorType", "solverType", "tolerance",
              "smootherWeight", "proOmega", "seed", "refLevel0NoPerm")

    def __init__(self, precision: int = 64, **params):
        self.L = lib()
        self.h = self.L.orc_create(precision)
        self.nv = 0
        self.nnz = 0
        self.set(**params)

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def set(self, **params):
        for k, v in params.items():
            if self.L.orc_set_param(self.h, k.encode(), float(v)) != 0:
                raise KeyError(k)
        return self

    # -- stage 1
    def pattern(self, nv, elems):
        elems = np.ascontiguousarray(elems, dtype=np.int32)
        self.nv = int(nv)
        self._elems = elems
        self.nnz = self.L.orc_pattern(self.h, nv, elems.shape[0], elems.shape[1], _p(elems))
        ptr = np.empty(nv + 1, dtype=np.int32)
        col = np.empty(self.nnz, dtype=np.int32)
        self.L.orc_get_pattern(self.h, _p(ptr), _p(col))
        return ptr, col

    def assemble(self, verts, labels=None, closed_form=False):
        verts = np.asarray(verts, dtype=np.float64)
        vx, vy = np.ascontiguousarray(verts[:, 0]), np.ascontiguousarray(verts[:, 1])
        e = self._elems
        if e.shape[1] == 4:
            vz = np.ascontiguousarray(verts[:, 2])
            lab = None if labels is None else np.ascontiguousarray(labels, dtype=np.int32)
            self.L.orc_assemble_tet(self.h, e.shape[0], _p(e), _p(vx), _p(vy), _p(vz),
                                    _p(lab) if lab is not None else None, int(closed_form))
        else:
            self.L.orc_assemble_tri(self.h, e.shape[0], _p(e), _p(vx), _p(vy), int(closed_form))
        return self.values()

    def values(self):
        v = np.empty(self.nnz, dtype=np.float64)
        self.L.orc_get_values(self.h, _p(v))
        return v

    def set_values(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert v.size == self.nnz
        self.L.orc_set_values(self.h, _p(v))

    def spmv(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        self.L.orc_spmv(self.h, _p(x), _p(y))
        return y

    # -- stage 2
    def setup(self):
        n = self.L.orc_setup(self.h)
        if n < 0:
            raise RuntimeError("oracle setup failed")
        return n

    def num_levels(self):
        return self.L.orc_num_levels(self.h)

    def level_rows(self, lev):
        return self.L.orc_level_rows(self.h, lev)

    def level_int(self, lev, name):
        n = self.L.orc_level_int(self.h, lev, name.encode(), None)
        if n < 0:
            raise KeyError(name)
        a = np.empty(n, dtype=np.int32)
        self.L.orc_level_int(self.h, lev, name.encode(), _p(a))
        return a

    def level_val(self, lev, name):
        n = self.L.orc_level_val(self.h, lev, name.encode(), None)
        if n < 0:
            raise KeyError(name)
        a = np.empty(n, dtype=np.float64)
        self.L.orc_level_val(self.h, lev, name.encode(), _p(a))
        return a

    # -- stage 3
    def solve(self, b, x0=None):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros_like(b) if x0 is None else np.array(x0, dtype=np.float64)
        it = self.L.orc_solve(self.h, _p(b), _p(x))
        if it < 0:
            raise RuntimeError("oracle solve before setup")
        return x, it

    def precond_permuted(self, r):
        r = np.ascontiguousarray(r, dtype=np.float64)
        z = np.empty_like(r)
        if self.L.orc_precond_permuted(self.h, _p(r), _p(z)) != 0:
            raise RuntimeError("oracle precond before setup")
        return z

    def final_relres(self):
        return self.L.orc_final_relres(self.h)

    def resid_history(self):
        n = self.L.orc_resid_history(self.h, None)
        a = np.empty(n, dtype=np.float64)
        self.L.orc_resid_history(self.h, _p(a))
        return a

    def time(self, what):
        return self.L.orc_time(self.h, what.encode())


def randomized_mis(xadj, adj, k, seed):
    xadj = np.ascontiguousarray(xadj, dtype=np.int32)
    adj = np.ascontiguousarray(adj, dtype=np.int32)
    mis = np.empty(xadj.size - 1, dtype=np.int32)
    lib().orc_randomized_mis(xadj.size - 1, _p(xadj), _p(adj), k, seed, _p(mis))
    return mis


def set_threads(t):
    lib().orc_set_threads(int(t))
