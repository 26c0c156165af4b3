"""Python host-side mirror of the reference's FEMSolver class over the C-ABI (ctypes).

Reference interface: class FEMSolver, /root/reference/src/FEMSolver.h:14-67 — same field names
(with the trailing underscore), same defaults (FEMSolver.cu:11-34), same call sequence:

    s = FEMSolver("mesh_base_name", isTetMesh=True)      # reads the mesh and assembles (ctor, :9-44)
    s.readMatlabSparseMatrix("A.mat")                    # optional (:177-356)
    s.solveFEM(x, b)                                     # AMG setup + solve (:58-93); overwrites x

Additive surface: `seed_`, `refLevel0NoPerm_`, in-memory meshes (`FEMSolver.from_arrays`), and the
setup()/solve() split.  All compute happens in libfemsolver_b200.so on the GPU; nothing here
falls back to the CPU: constructing a solver without the library or without a device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import meshio

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libfemsolver_b200.so")
_lib = None


class FEMSolverError(RuntimeError):
    pass


def load_library():
    """Loads the CUDA library; fails loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIBPATH):
        raise FEMSolverError(f"{_LIBPATH} is missing: run `python sci-solver_fem_b200/build.py` "
                             "(the CUDA extension is the only implementation of this path)")
    L = C.CDLL(_LIBPATH)
    vp, ci, cd, cll, cs = C.c_void_p, C.c_int, C.c_double, C.c_longlong, C.c_char_p
    sig = {
        "fsb_version": (ci, []), "fsb_device_count": (ci, []),
        "fsb_create": (ci, [C.POINTER(vp), ci]), "fsb_destroy": (None, [vp]), "fsb_last_error": (cs, [vp]),
        "fsb_set_param": (ci, [vp, cs, cd]), "fsb_get_param": (ci, [vp, cs, C.POINTER(cd)]),
        "fsb_set_tet_mesh": (ci, [vp, ci, vp, ci, vp, vp]), "fsb_set_tri_mesh": (ci, [vp, ci, vp, ci, vp]),
        "fsb_set_tet_mesh_device": (ci, [vp, ci, vp, ci, vp, vp]), "fsb_set_tri_mesh_device": (ci, [vp, ci, vp, ci, vp]),
        "fsb_assemble": (ci, [vp]), "fsb_matrix_rows": (ci, [vp]), "fsb_matrix_nnz": (cll, [vp]),
        "fsb_get_matrix_csr": (ci, [vp, vp, vp, vp]), "fsb_set_matrix_values": (ci, [vp, vp]),
        "fsb_set_matrix_csr": (ci, [vp, ci, cll, vp, vp, vp]),
        "fsb_setup": (ci, [vp]), "fsb_num_levels": (ci, [vp]), "fsb_level_rows": (ci, [vp, ci]), "fsb_level_nnz": (cll, [vp, ci]),
        "fsb_level_stat": (cll, [vp, ci, cs]), "fsb_level_int": (cll, [vp, ci, cs, vp, cll]), "fsb_level_val": (cll, [vp, ci, cs, vp, cll]),
        "fsb_solve": (ci, [vp, vp, vp, C.POINTER(ci), C.POINTER(cd)]),
        "fsb_solve_device": (ci, [vp, vp, vp, C.POINTER(ci), C.POINTER(cd)]),
        "fsb_solve_fem": (ci, [vp, vp, vp, C.POINTER(ci), C.POINTER(cd)]),
        "fsb_resid_history": (ci, [vp, vp, ci]),
        "fsb_spmv_fine_device": (ci, [vp, vp, vp]), "fsb_apply_matrix_device": (ci, [vp, vp, vp]), "fsb_apply_matrix": (ci, [vp, vp, vp]), "fsb_precondition_device": (ci, [vp, vp, vp]),
        "fsb_time_ms": (cd, [vp, cs]), "fsb_last_launches": (cll, [vp]), "fsb_stream": (vp, [vp]),
        "fsb_profile_report": (ci, [vp, vp, ci]),
        "fsb_dist_prepare": (ci, [vp, ci, ci]), "fsb_dist_blob": (ci, [vp, vp, C.POINTER(cll)]), "fsb_dist_blob_bytes": (ci, []),
        "fsb_dist_connect": (ci, [vp, vp]), "fsb_dist_disconnect": (ci, [vp]), "fsb_dist_ranges": (ci, [vp, vp, vp, vp]),
        "fsb_dist_level_ranges": (ci, [vp, ci, vp, vp, vp]), "fsb_dist_info": (ci, [vp, vp, vp, vp, vp]), "fsb_dist_interior": (ci, [vp, ci, vp]),
        "fsb_split_by_weight": (None, [ci, vp, ci, vp]),
        "fsb_tet_mass_integrals": (None, [vp]), "fsb_tri_quadrature": (None, [vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = (
    "fsb_version fsb_device_count fsb_create fsb_destroy fsb_last_error fsb_set_param fsb_get_param fsb_set_tet_mesh "
    "fsb_set_tri_mesh fsb_set_tet_mesh_device fsb_set_tri_mesh_device fsb_assemble fsb_matrix_rows fsb_matrix_nnz "
    "fsb_get_matrix_csr fsb_set_matrix_values fsb_set_matrix_csr fsb_setup fsb_num_levels fsb_level_rows fsb_level_nnz "
    "fsb_level_stat fsb_level_int fsb_level_val fsb_solve fsb_solve_device fsb_solve_fem fsb_resid_history fsb_spmv_fine_device "
    "fsb_precondition_device fsb_time_ms fsb_last_launches fsb_stream fsb_tet_mass_integrals fsb_tri_quadrature "
    "fsb_profile_report fsb_dist_prepare fsb_dist_blob fsb_dist_blob_bytes fsb_dist_connect fsb_dist_disconnect fsb_dist_ranges fsb_dist_level_ranges fsb_dist_info fsb_dist_interior "
    "fsb_split_by_weight fsb_apply_matrix_device fsb_apply_matrix").split()


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def split_by_weight(weights, nranks):
    """Contiguous, weight-balanced split (host helper of the sharded solve): first item of every rank."""
    w = np.ascontiguousarray(weights, dtype=np.int64)
    out = np.zeros(nranks + 1, dtype=np.int32)
    load_library().fsb_split_by_weight(w.size, _p(w), int(nranks), _p(out))
    return out


def exchange_handles_torch(payload: bytes, group=None):
    """All-gathers one fixed-size byte string per rank with torch.distributed (any backend)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    outs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(outs, mine, group=group)
    return [o.cpu().numpy().tobytes() for o in outs]


_FIELDS = {  # reference field -> (C-ABI parameter, default)   FEMSolver.cu:11-34
    "verbose_": ("verbose", False), "maxLevels_": ("maxLevels", 100), "maxIters_": ("maxIters", 100),
    "preInnerIters_": ("preInnerIters", 5), "postInnerIters_": ("postInnerIters", 5), "postRelaxes_": ("postRelaxes", 1),
    "cycleIters_": ("cycleIters", 1), "dsType_": ("dsType", 0), "topSize_": ("topSize", 256),
    "randMisParameters_": ("randMisParameters", 90102), "partitionMaxSize_": ("partitionMaxSize", 512),
    "aggregatorType_": ("aggregatorType", 0), "convergeType_": ("convergeType", 0), "tolerance_": ("tolerance", 1e-6),
    "cycleType_": ("cycleType", 0), "solverType_": ("solverType", 0), "smootherWeight_": ("smootherWeight", 1.0),
    "proOmega_": ("proOmega", 0.67), "device_": ("device", 0), "blockSize_": ("blockSize", 256),
    # additive
    "seed_": ("seed", 0), "refLevel0NoPerm_": ("refLevel0NoPerm", 0), "useGraphs_": ("useGraphs", 1), "checkEvery_": ("checkEvery", 2),
    "profile_": ("profile", 0),
}


class FEMSolver:
    def __init__(self, fname: str | None = "../src/test/test_data/simple", isTetMesh: bool = True, verbose: bool = False,
                 device: int = 0):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.fsb_create(C.byref(h), int(device))
        if rc != 0:
            raise FEMSolverError("fsb_create failed: " + (self._L.fsb_last_error(None) or b"").decode())
        self._h = h
        for f, (_, default) in _FIELDS.items():
            object.__setattr__(self, f, default)
        self.verbose_ = bool(verbose)
        self.device_ = int(device)
        self.filename_ = fname
        self.vertices = None
        self.elements = None
        self.matlabels = None
        self.iterations = 0
        self.relres = -1.0
        if fname is not None:
            if isTetMesh:
                v, t, lab = meshio.read_node_ele(fname)
                self._set_mesh(v, t, lab)
            else:
                v, f = meshio.read_trimesh(fname)
                self._set_mesh(v, f, None)
            self.getMatrixFromMesh()

    @classmethod
    def from_arrays(cls, vertices, elements, matlabels=None, verbose=False, device=0):
        """In-memory mesh (synthetic cubes): same as the file constructor without the parse."""
        s = cls(None, verbose=verbose, device=device)
        s._set_mesh(vertices, elements, matlabels)
        s.getMatrixFromMesh()
        return s

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.fsb_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc != 0:
            msg = (self._L.fsb_last_error(self._h) or b"").decode()
            if rc == -1:
                raise ValueError(msg)  # std::invalid_argument upstream
            raise FEMSolverError(msg)

    def _push_params(self):
        for f, (name, _) in _FIELDS.items():
            self._check(self._L.fsb_set_param(self._h, name.encode(), float(getattr(self, f))))

    def _set_mesh(self, vertices, elements, matlabels):
        v = np.ascontiguousarray(vertices, dtype=np.float64)
        if v.shape[1] == 2:
            v = np.concatenate([v, np.zeros((v.shape[0], 1))], axis=1)
        e = np.ascontiguousarray(elements, dtype=np.int32)
        self.vertices, self.elements = v, e
        if e.shape[1] == 4:
            lab = None if matlabels is None else np.ascontiguousarray(matlabels, dtype=np.int32)
            self.matlabels = lab
            self._check(self._L.fsb_set_tet_mesh(self._h, v.shape[0], _p(v), e.shape[0], _p(e), _p(lab)))
        else:
            self._check(self._L.fsb_set_tri_mesh(self._h, v.shape[0], _p(v), e.shape[0], _p(e)))

    # ------------------------------------------------------------------ reference API
    def getMatrixFromMesh(self):
        self._push_params()
        self._check(self._L.fsb_assemble(self._h))

    def getMatrixRows(self):
        return self._L.fsb_matrix_rows(self._h)

    def checkMatrixForValidContents(self):
        if self.getMatrixRows() == 0:
            raise ValueError("Error no matrix specified")

    def matrix_csr(self):
        n, nnz = self.getMatrixRows(), self._L.fsb_matrix_nnz(self._h)
        ptr = np.empty(n + 1, dtype=np.int32); col = np.empty(nnz, dtype=np.int32); val = np.empty(nnz, dtype=np.float64)
        self._check(self._L.fsb_get_matrix_csr(self._h, _p(ptr), _p(col), _p(val)))
        return ptr, col, val

    def set_matrix_values(self, val):
        val = np.ascontiguousarray(val, dtype=np.float64)
        assert val.size == self._L.fsb_matrix_nnz(self._h)
        self._check(self._L.fsb_set_matrix_values(self._h, _p(val)))

    def readMatlabSparseMatrix(self, filename: str, float_round: bool = True) -> int:
        """FEMSolver::readMatlabSparseMatrix (FEMSolver.cu:177-356): the assembled values are replaced
        by 1e-12 on the mesh pattern and the file matrix is added on top (pattern union); values pass
        through float as upstream (A_h_ is an ell_matrix<int,float>) unless float_round=False."""
        try:
            nr, nc, jc, ir, pr = meshio.read_mat_sparse(filename)
        except (OSError, ValueError) as e:
            import sys
            print(str(e), file=sys.stderr)
            return 1
        fptr, fcol, fval = meshio.csc_to_csr(nr, nc, jc, ir, pr)
        self.set_matrix_from_csr(fptr, fcol, fval, float_round=float_round)
        return 0

    def set_matrix_from_csr(self, fptr, fcol, fval, float_round: bool = True):
        import scipy.sparse as sp
        ptr, col, val = self.matrix_csr()
        n = ptr.size - 1
        if float_round:
            fval = fval.astype(np.float32).astype(np.float64)
            base = np.full(val.size, np.float64(np.float32(1e-12)))
        else:
            base = np.full(val.size, 1e-12)
        M = sp.csr_matrix((base, col, ptr), shape=(n, n)) + sp.csr_matrix((fval, fcol, fptr), shape=(n, n))
        if float_round:
            M.data = M.data.astype(np.float32).astype(np.float64)
        M.sort_indices()
        p2 = np.ascontiguousarray(M.indptr, dtype=np.int32); c2 = np.ascontiguousarray(M.indices, dtype=np.int32)
        v2 = np.ascontiguousarray(M.data, dtype=np.float64)
        self._check(self._L.fsb_set_matrix_csr(self._h, n, v2.size, _p(p2), _p(c2), _p(v2)))

    @staticmethod
    def readMatlabArray(filename: str):
        return meshio.read_mat_array(filename)

    @staticmethod
    def writeMatlabArray(filename: str, array) -> int:
        meshio.write_mat_array(filename, array)
        return 0

    def setup(self):
        self.checkMatrixForValidContents()
        self._push_params()
        self._check(self._L.fsb_setup(self._h))

    def solve(self, x, b):
        """x: initial guess in, solution out (modified in place and returned)."""
        b = np.ascontiguousarray(b, dtype=np.float64)
        assert isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous
        self._push_params()
        it, rr = C.c_int(0), C.c_double(0)
        self._check(self._L.fsb_solve(self._h, _p(b), _p(x), C.byref(it), C.byref(rr)))
        self.iterations, self.relres = it.value, rr.value
        return x

    def solveFEM(self, x, b):
        """FEMSolver::solveFEM (FEMSolver.cu:58-93): rebuilds the hierarchy, then solves."""
        self.setup()
        return self.solve(x, b)

    # ------------------------------------------------------------------ introspection
    def num_levels(self):
        return self._L.fsb_num_levels(self._h)

    def level_rows(self, lev):
        return self._L.fsb_level_rows(self._h, lev)

    def level_nnz(self, lev):
        return self._L.fsb_level_nnz(self._h, lev)

    def level_int(self, lev, name):
        n = self._L.fsb_level_int(self._h, lev, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        a = np.empty(n, dtype=np.int32)
        if n:
            self._L.fsb_level_int(self._h, lev, name.encode(), _p(a), n)
        return a

    def level_val(self, lev, name):
        n = self._L.fsb_level_val(self._h, lev, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        a = np.empty(n, dtype=np.float64)
        if n:
            self._L.fsb_level_val(self._h, lev, name.encode(), _p(a), n)
        return a

    def resid_history(self):
        n = self._L.fsb_resid_history(self._h, None, 0)
        a = np.empty(max(n, 0), dtype=np.float64)
        if n > 0:
            self._L.fsb_resid_history(self._h, _p(a), n)
        return a

    def time_ms(self, stage):
        return self._L.fsb_time_ms(self._h, stage.encode())

    def last_launches(self):
        return self._L.fsb_last_launches(self._h)

    def profile_report(self):
        """{(kernel, level): (launches, total_ms)} of the last solve run with profile_ = 1."""
        n = self._L.fsb_profile_report(self._h, None, 0)
        buf = C.create_string_buffer(max(n, 1))
        self._L.fsb_profile_report(self._h, buf, n)
        out = {}
        for ln in buf.value.decode().splitlines():
            name, lev, cnt, ms = ln.split()
            out[(name, int(lev))] = (int(cnt), float(ms))
        return out

    # ------------------------------------------------------------------ stage 4: sharded solve
    def level_stat(self, level: int, name: str) -> int:
        return int(self._L.fsb_level_stat(self._h, int(level), name.encode()))

    def dist_connect(self, rank: int, world: int, allgather):
        """Switches the PCG solve to the sharded mode.  `allgather(bytes) -> list[bytes]` exchanges the ranks'
        connection records (IPC handle + receive-buffer layout) in rank order (see exchange_handles_torch)."""
        self._check(self._L.fsb_dist_prepare(self._h, int(rank), int(world)))
        if world == 1:
            return
        nb = self._L.fsb_dist_blob_bytes()
        buf = C.create_string_buffer(nb)
        nbytes = C.c_longlong(0)
        self._check(self._L.fsb_dist_blob(self._h, buf, C.byref(nbytes)))
        blobs = allgather(buf.raw)
        assert len(blobs) == world and all(len(h) == nb for h in blobs)
        assert blobs[rank] == buf.raw, "all-gather returned the records in the wrong order"
        self._check(self._L.fsb_dist_connect(self._h, C.create_string_buffer(b"".join(blobs), nb * world)))

    def dist_disconnect(self):
        self._check(self._L.fsb_dist_disconnect(self._h))

    def dist_ranges(self, level: int = 0):
        pb, rb, ab = (np.zeros(9, dtype=np.int32) for _ in range(3))
        n = self._L.fsb_dist_level_ranges(self._h, int(level), _p(pb), _p(rb), _p(ab))
        if n < 0:
            raise ValueError(f"level {level} is not sharded")
        return pb[: n + 1].copy(), rb[: n + 1].copy(), ab[: n + 1].copy()

    def dist_info(self):
        """{'sharded_levels', 'user_range' (host-copy slice of a sharded solve), 'halo_values' per sharded level}."""
        ns, lo, hi = C.c_int(0), C.c_int(0), C.c_int(0)
        hv = np.zeros(4 * 16, dtype=np.int64)
        self._L.fsb_dist_info(self._h, C.byref(ns), C.byref(lo), C.byref(hi), _p(hv))
        interior = []
        for l in range(ns.value):
            o = np.zeros(6, dtype=np.int32)
            self._L.fsb_dist_interior(self._h, l, _p(o))
            interior.append({"operator_rows": [int(o[0]), int(o[1])], "restriction_rows": [int(o[2]), int(o[3])], "prolongator_rows": [int(o[4]), int(o[5])]})
        return {"sharded_levels": ns.value, "user_range": (lo.value, hi.value),
                "halo_values": [dict(zip(("operator", "residual", "down", "up"), map(int, hv[4 * l: 4 * l + 4]))) for l in range(ns.value)],
                "interior_ranges": interior}

    def apply_matrix(self, x):
        """y = A x with the assembled (user-ordered) matrix; host vectors, the product runs on the GPU."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        self._check(self._L.fsb_apply_matrix(self._h, _p(x), _p(y)))
        return y

    def apply_matrix_device(self, x_ptr: int, y_ptr: int):
        """y = A x on the device (user ordering)."""
        self._check(self._L.fsb_apply_matrix_device(self._h, C.c_void_p(x_ptr), C.c_void_p(y_ptr)))

    def solve_device(self, x_ptr: int, b_ptr: int):
        """Device-pointer variant of solve(): b and x are already resident in HBM."""
        self._push_params()
        it, rr = C.c_int(0), C.c_double(0)
        self._check(self._L.fsb_solve_device(self._h, C.c_void_p(b_ptr), C.c_void_p(x_ptr), C.byref(it), C.byref(rr)))
        self.iterations, self.relres = it.value, rr.value

    # raw handle for bench.py (device-pointer entry points)
    @property
    def handle(self):
        return self._h

    @property
    def lib(self):
        return self._L
