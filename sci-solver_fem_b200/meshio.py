"""Host-side data formats either side of the solve path (numpy only, no torch, no oracle).

Mirrors what the reference reads/writes around FEMSolver:
  * TetGen .node/.ele      -> reference src/core/cuda/tetmesh.cu:251-376 (TetMesh::read)
  * ASCII PLY              -> reference src/core/aggmis/cuda/TriMesh_io.cu:259,874-878
  * MATLAB v5 (-v6) .mat   -> reference src/FEMSolver.cu:177-356 (sparse), :358-449 (array), :451-517 (writer)
  * synthetic Kuhn cubes   -> SURVEY.md section 4 / section 8(d): the generator of
                              CubeMesh_size256step16_correct and of the BASELINE.json configs.
"""
from __future__ import annotations

import struct

import numpy as np

# per-cell split of test_data/CubeMesh_size256step16_correct.ele (corner bits are x,y,z)
_KUHN = (
    ("000", "101", "011", "001"),
    ("100", "010", "000", "101"),
    ("000", "011", "101", "010"),
    ("101", "011", "111", "010"),
    ("010", "101", "100", "111"),
    ("100", "110", "010", "111"),
)


def kuhn_cube(N: int, h: float | None = None, scale=(1.0, 1.0, 1.0), float_round: bool = True):
    """Structured tet cube with N cells per side (6 tets per cell, x-fastest cells).

    Returns (vertices[nv,3] float64, tets[ne,4] int32).  Vertex id = x + (N+1) y + (N+1)^2 z,
    coordinates h*(x,y,z) (h = 1/N by default), rounded through float32 when `float_round`
    to mimic the reference's `%f`-into-float parse (tetmesh.cu:285-292).
    """
    if h is None:
        h = 1.0 / N
    n1 = N + 1
    g = np.arange(n1, dtype=np.float64) * h
    Z, Y, X = np.meshgrid(g * scale[2], g * scale[1], g * scale[0], indexing="ij")
    verts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    if float_round:
        verts = verts.astype(np.float32).astype(np.float64)
    c = np.arange(N, dtype=np.int64)
    CZ, CY, CX = np.meshgrid(c, c, c, indexing="ij")
    base = (CX + n1 * CY + n1 * n1 * CZ).ravel()  # x-fastest
    tets = np.empty((base.size, 6, 4), dtype=np.int32)
    for t, corners in enumerate(_KUHN):
        for k, bits in enumerate(corners):
            off = int(bits[0]) + n1 * int(bits[1]) + n1 * n1 * int(bits[2])
            tets[:, t, k] = base + off
    return verts, tets.reshape(-1, 4)


def kuhn_cell_labels(N: int, block: int = 8):
    """BASELINE config 5(i): matlabel = 1 + ((cx/block + cy/block + cz/block) mod 6) per cell, 6 tets each."""
    c = np.arange(N, dtype=np.int64) // block
    CZ, CY, CX = np.meshgrid(c, c, c, indexing="ij")
    lab = (1 + (CX + CY + CZ) % 6).ravel().astype(np.int32)
    return np.repeat(lab, 6)


def grid_tri(nx: int, ny: int, lx: float = 1.0, ly: float = 1.0):
    """Structured triangle grid ((nx+1)(ny+1) vertices, 2 nx ny triangles), z = 0."""
    gx = np.linspace(0.0, lx, nx + 1)
    gy = np.linspace(0.0, ly, ny + 1)
    Y, X = np.meshgrid(gy, gx, indexing="ij")
    verts = np.stack([X.ravel(), Y.ravel(), np.zeros(X.size)], axis=1)
    i = np.arange(nx)
    j = np.arange(ny)
    J, I = np.meshgrid(j, i, indexing="ij")
    v0 = (I + (nx + 1) * J).ravel()
    t1 = np.stack([v0, v0 + 1, v0 + nx + 2], axis=1)
    t2 = np.stack([v0, v0 + nx + 2, v0 + nx + 1], axis=1)
    tris = np.empty((2 * v0.size, 3), dtype=np.int32)
    tris[0::2] = t1
    tris[1::2] = t2
    return verts, tris


# ----------------------------------------------------------------------------- TetGen
def read_node_ele(base: str):
    """TetMesh::read (tetmesh.cu:251-376): float-rounded coordinates, 0/1-based auto-detect,
    optional material column, and the second `minidx == 1` decrement (:358-372)."""
    def lines(path):
        with open(path) as f:
            for ln in f:
                ln = ln.strip()
                if not ln or ln[0] == "#":
                    continue
                yield ln

    it = lines(base + ".node")
    nv = int(next(it).split()[0])
    verts = np.zeros((nv, 3), dtype=np.float64)
    i = 0
    for ln in it:
        p = ln.split()
        if i >= nv:
            break
        verts[i] = [np.float32(p[1]), np.float32(p[2]), np.float32(p[3])]
        i += 1
    it = lines(base + ".ele")
    hdr = next(it).split()
    ne, haslabel = int(hdr[0]), int(hdr[2])
    tets = np.zeros((ne, 4), dtype=np.int32)
    labels = np.zeros(ne, dtype=np.int32)
    i = 0
    for ln in it:
        p = ln.split()
        if i >= ne:
            break
        tets[i] = [int(p[1]), int(p[2]), int(p[3]), int(p[4])]
        if haslabel != 0:
            labels[i] = int(p[5])
        i += 1
    if not (tets == 0).any():
        tets -= 1
    if tets.min() == 1:
        tets -= 1
    return verts, tets, labels


def write_node_ele(base: str, verts, tets, labels=None):
    with open(base + ".node", "w") as f:
        f.write(f"{len(verts)} 3 0 0\n")
        for i, v in enumerate(verts):
            f.write(f"{i} {v[0]:.10f} {v[1]:.10f} {v[2]:.10f}\n")
    with open(base + ".ele", "w") as f:
        f.write(f"{len(tets)} 4 {0 if labels is None else 1}\n")
        for i, t in enumerate(tets):
            if labels is None:
                f.write(f"{i} {t[0]} {t[1]} {t[2]} {t[3]}\n")
            else:
                f.write(f"{i} {t[0]} {t[1]} {t[2]} {t[3]} {labels[i]}\n")


# ----------------------------------------------------------------------------- PLY
def read_ply_ascii(path: str):
    """ASCII PLY with x y z first per vertex and `n i j k` faces (%lf parse, TriMesh_io.cu:874-878)."""
    with open(path) as f:
        assert f.readline().strip() == "ply"
        nv = nf = 0
        nprop = 0
        in_vertex = False
        while True:
            ln = f.readline().strip()
            if ln.startswith("format"):
                assert "ascii" in ln, "only ASCII PLY is supported"
            elif ln.startswith("element vertex"):
                nv = int(ln.split()[2]); in_vertex = True
            elif ln.startswith("element face"):
                nf = int(ln.split()[2]); in_vertex = False
            elif ln.startswith("element"):
                in_vertex = False
            elif ln.startswith("property") and in_vertex:
                nprop += 1
            elif ln == "end_header":
                break
        toks = f.read().split()
    verts = np.array(toks[: nv * nprop], dtype=np.float64).reshape(nv, nprop)[:, :3].copy()
    rest = toks[nv * nprop:]
    tris = np.zeros((nf, 3), dtype=np.int32)
    p = 0
    for i in range(nf):
        k = int(rest[p])
        assert k == 3, "only triangles"
        tris[i] = [int(rest[p + 1]), int(rest[p + 2]), int(rest[p + 3])]
        p += 1 + k
    return verts, tris


def write_ply_ascii(path: str, verts, tris):
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\n")
        f.write(f"element vertex {len(verts)}\nproperty float x\nproperty float y\nproperty float z\n")
        f.write(f"element face {len(tris)}\nproperty list uchar int vertex_indices\nend_header\n")
        for v in verts:
            f.write(f"{v[0]:.17g} {v[1]:.17g} {v[2]:.17g}\n")
        for t in tris:
            f.write(f"3 {t[0]} {t[1]} {t[2]}\n")


# ----------------------------------------------------------------------------- MATLAB v5
def _read_name(buf, off):
    """array-name element: small-data form or long form (FEMSolver.cu:232-256)."""
    t, ln = struct.unpack_from("<HH", buf, off)
    off += 4
    align = 4
    if ln == 0:
        (ln,) = struct.unpack_from("<I", buf, off)
        off += 4
        align = 8
    if t not in (1, 2):
        raise ValueError("invalid array-name type %d" % t)
    if ln % align:
        ln += align - ln % align
    return off + ln


def read_mat_sparse(path: str):
    """Returns (nrows, ncols, jc[ncols+1], ir[nnz], pr[nnz]) — the CSC arrays of a v5 sparse matrix."""
    buf = open(path, "rb").read()
    off = 128
    (t,) = struct.unpack_from("<i", buf, off)
    if t == 15:
        raise ValueError("Compression not supported. Save matlab data with '-v6' option.")
    if t != 14:
        raise ValueError("not a matlab matrix")
    off += 8  # type, size
    t, nb, flags, nzmax = struct.unpack_from("<iiII", buf, off)
    off += 16
    if (flags & 0xFF) != 5:
        raise ValueError("This is not a sparse matrix file.")
    t, nb, xd, yd = struct.unpack_from("<iiii", buf, off)
    off += 16
    off = _read_name(buf, off)

    def block(off, dtype, ok):
        t, nb = struct.unpack_from("<ii", buf, off)
        if t not in ok:
            raise ValueError("unexpected element type %d" % t)
        off += 8
        a = np.frombuffer(buf, dtype=dtype, count=nb // np.dtype(dtype).itemsize, offset=off).copy()
        off += nb + (nb % 8)  # upstream skips `bytes % 8` (FEMSolver.cu:270), exact for int32 blocks
        return a, off

    ir, off = block(off, "<i4", (5, 6))
    jc, off = block(off, "<i4", (5, 6))
    pr, off = block(off, "<f8", (9,))
    nnz = int(jc[yd])
    return xd, yd, jc[: yd + 1].copy(), ir[:nnz].copy(), pr[:nnz].copy()


def read_mat_array(path: str):
    buf = open(path, "rb").read()
    off = 128
    (t,) = struct.unpack_from("<i", buf, off)
    if t != 14:
        raise ValueError("not a matlab matrix")
    off += 8
    t, nb, flags, nzmax = struct.unpack_from("<iiII", buf, off)
    off += 16
    if (flags & 0xFF) == 5:
        raise ValueError("This import routine is not for a sparse matrix file.")
    off += 16  # dims
    t, ln = struct.unpack_from("<HH", buf, off)
    off += 4
    if ln % 4:
        ln += 4 - ln % 4
    off += ln
    t, nb = struct.unpack_from("<iI", buf, off)
    if t != 9:
        raise ValueError("Matrix data type must be miDOUBLE")
    off += 8
    return np.frombuffer(buf, dtype="<f8", count=nb // 8, offset=off).copy()


def write_mat_array(path: str, arr, name: str = "x_h"):
    """Same container layout as FEMSolver::writeMatlabArray (FEMSolver.cu:451-517): one n x 1 double matrix."""
    arr = np.ascontiguousarray(arr, dtype="<f8")
    desc = b"MATLAB 5.0 MAT-file, Platform: GLNXA64, Created by SCI-Solver_FEM."
    hdr = desc.ljust(116, b" ") + b"\0" * 8 + struct.pack("<H", 0x0100) + b"IM"
    nm = name.encode()[:4]
    body = struct.pack("<iiII", 6, 8, 6, 0)
    body += struct.pack("<iiii", 5, 8, arr.size, 1)
    body += struct.pack("<HH", 1, len(nm)) + nm.ljust(4, b"\0")
    body += struct.pack("<iI", 9, arr.size * 8) + arr.tobytes()
    with open(path, "wb") as f:
        f.write(hdr + struct.pack("<iI", 14, len(body)) + body)


def write_mat_sparse(path: str, nrows: int, ncols: int, jc, ir, pr, name: str = "A"):
    """Uncompressed v5 sparse matrix (mxSPARSE_CLASS, int32 ir/jc, double pr) that
    FEMSolver::readMatlabSparseMatrix (FEMSolver.cu:177-356) accepts."""
    jc = np.ascontiguousarray(jc, dtype="<i4"); ir = np.ascontiguousarray(ir, dtype="<i4"); pr = np.ascontiguousarray(pr, dtype="<f8")
    desc = b"MATLAB 5.0 MAT-file, Platform: GLNXA64, Created by SCI-Solver_FEM."
    hdr = desc.ljust(116, b" ") + b"\0" * 8 + struct.pack("<H", 0x0100) + b"IM"

    def pad8(b):
        return b + b"\0" * ((8 - len(b) % 8) % 8)

    nm = name.encode()[:4]
    body = struct.pack("<iiII", 6, 8, 5, ir.size)
    body += struct.pack("<iiii", 5, 8, nrows, ncols)
    body += struct.pack("<HH", 1, len(nm)) + nm.ljust(4, b"\0")
    body += struct.pack("<ii", 5, ir.size * 4) + pad8(ir.tobytes())
    body += struct.pack("<ii", 5, jc.size * 4) + pad8(jc.tobytes())
    body += struct.pack("<ii", 9, pr.size * 8) + pr.tobytes()
    with open(path, "wb") as f:
        f.write(hdr + struct.pack("<iI", 14, len(body)) + body)


def csc_to_csr(nrows, ncols, jc, ir, pr):
    """CSC -> CSR with ascending columns (the reference sorts entries by (row, col), FEMSolver.cu:301)."""
    nnz = ir.size
    cols = np.repeat(np.arange(ncols, dtype=np.int32), np.diff(jc))
    order = np.lexsort((cols, ir))
    rows = ir[order]
    ptr = np.zeros(nrows + 1, dtype=np.int32)
    np.add.at(ptr, rows + 1, 1)
    ptr = np.cumsum(ptr).astype(np.int32)
    return ptr, cols[order].astype(np.int32), pr[order].astype(np.float64)
