"""Host-side data formats either side of the solve path (numpy only, no torch, no oracle).

Mirrors what the reference reads/writes around FEMSolver:
  * TetGen .node/.ele      -> reference src/core/cuda/tetmesh.cu:251-376 (TetMesh::read)
  * triangle meshes        -> reference src/core/aggmis/cuda/TriMesh_io.cu:146-1276 (PLY, 3DS, VVD, RAY, OBJ, OFF, SM)
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
# ------------------------------------------------------------------------------------- triangle meshes
# TriMesh::read (aggmis/cuda/TriMesh_io.cu:146-256) recognises the format from the first bytes of the file.
# Covered here: PLY (ascii, binary little/big endian; any scalar property types, the vertex_indices list with
# any integer count/index types, face lists or triangle strips; other elements — also range_grid — are skipped), 3DS,
# VVD, RAY, OBJ, OFF and old-style SM.  Polygons are cut into triangles by upstream's rule (tess, :1239-1270).
_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
              "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
              "double": "f8", "float64": "f8"}


def tessellate(verts, face):
    """Triangles of one polygon (vertex index list): 3 -> itself; 4 -> split along the shorter diagonal
    (0-2 if strictly shorter, else 1-3); 5 and more -> fan around the first vertex; fewer than 3 -> nothing."""
    k = len(face)
    if k < 3:
        return []
    if k == 3:
        return [tuple(face)]
    if k == 4:
        p = [np.asarray(verts[i], dtype=np.float64) for i in face]
        d02 = float(np.sum((p[0] - p[2]) ** 2))
        d13 = float(np.sum((p[1] - p[3]) ** 2))
        i = 0 if d02 < d13 else 1
        return [(face[i], face[(i + 1) % 4], face[(i + 2) % 4]), (face[i], face[(i + 2) % 4], face[(i + 3) % 4])]
    return [(face[0], face[i - 1], face[i]) for i in range(2, k)]


def _finish_trimesh(verts, polys):
    verts = np.ascontiguousarray(verts, dtype=np.float64).reshape(-1, 3)
    tris = [t for f in polys for t in tessellate(verts, f)]
    tris = np.asarray(tris, dtype=np.int32).reshape(-1, 3)
    if len(verts) == 0 and len(tris) == 0:
        raise ValueError("empty mesh")
    if len(tris) and (tris.min() < 0 or tris.max() >= len(verts)):
        raise ValueError("face index out of range")
    return verts, tris


def unpack_tstrips(idx):
    """Triangle strips as PLY stores them (`element tristrips`, one index list, strips separated by -1) -> triangles, the
    orientation flipped on every second triangle of a strip.  Upstream reads the list (TriMesh_io.cu:1021-1063) but never
    unpacks it in this code base (`need_faces()` / `convert_strips` are commented out, :478), i.e. a strip file gives it a
    mesh without faces; here the strips become the faces they encode."""
    tris, run = [], []
    for v in list(idx) + [-1]:
        if v < 0:
            for i in range(len(run) - 2):
                a, b, c = run[i], run[i + 1], run[i + 2]
                if a != b and b != c and a != c:  # degenerate triangles only stitch strips together
                    tris.append([a, b, c] if i % 2 == 0 else [b, a, c])
            run = []
        else:
            run.append(int(v))
    return tris


def read_ply(path: str):
    """PLY, ascii or binary (either endianness); faces as `element face` lists or as `element tristrips`; a `range_grid`
    element is skipped as upstream does (its reader is commented out, TriMesh_io.cu:1066-1108)."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.find(b"end_header")
    if not data.startswith(b"ply") or end < 0:
        raise ValueError("not a PLY file")
    nl = data.find(b"\n", end)
    header = data[:end].decode("ascii", "replace").splitlines()
    body = data[nl + 1:]
    fmt, elements = None, []          # elements: [name, count, [(kind, name, types...)]]
    for ln in header[1:]:
        t = ln.split()
        if not t or t[0] in ("comment", "obj_info"):
            continue
        if t[0] == "format":
            fmt = t[1]
        elif t[0] == "element":
            elements.append([t[1], int(t[2]), []])
        elif t[0] == "property" and elements:
            if t[1] == "list":
                elements[-1][2].append(("list", t[4], _PLY_TYPES[t[2]], _PLY_TYPES[t[3]]))
            else:
                elements[-1][2].append(("scalar", t[2], _PLY_TYPES[t[1]]))
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        raise ValueError("unknown PLY format")
    verts, polys = np.zeros((0, 3)), []
    if fmt == "ascii":
        toks = body.split()
        pos = 0
        for name, count, props in elements:
            rows = []
            for _ in range(count):
                row = {}
                for pr in props:
                    if pr[0] == "scalar":
                        row[pr[1]] = float(toks[pos]); pos += 1
                    else:
                        k = int(float(toks[pos])); pos += 1
                        row[pr[1]] = [int(float(x)) for x in toks[pos:pos + k]]; pos += k
                rows.append(row)
            if name == "vertex":
                verts = np.array([[r["x"], r["y"], r["z"]] for r in rows], dtype=np.float64).reshape(-1, 3)
            elif name == "face":
                key = "vertex_indices" if (rows and "vertex_indices" in rows[0]) else "vertex_index"
                polys = [r[key] for r in rows]
            elif name == "tristrips":
                polys = [tri for r in rows for tri in unpack_tstrips(r["vertex_indices"])]
    else:
        e = "<" if fmt == "binary_little_endian" else ">"
        mv = memoryview(body)
        pos = 0
        for name, count, props in elements:
            if all(pr[0] == "scalar" for pr in props):
                dt = np.dtype([(pr[1], e + pr[2]) for pr in props])
                arr = np.frombuffer(mv, dtype=dt, count=count, offset=pos)
                pos += dt.itemsize * count
                if name == "vertex":
                    verts = np.stack([arr["x"], arr["y"], arr["z"]], axis=1).astype(np.float64)
                continue
            rows = []
            for _ in range(count):
                row = {}
                for pr in props:
                    if pr[0] == "scalar":
                        dt = np.dtype(e + pr[2])
                        row[pr[1]] = np.frombuffer(mv, dtype=dt, count=1, offset=pos)[0]; pos += dt.itemsize
                    else:
                        ct, it = np.dtype(e + pr[2]), np.dtype(e + pr[3])
                        k = int(np.frombuffer(mv, dtype=ct, count=1, offset=pos)[0]); pos += ct.itemsize
                        row[pr[1]] = [int(x) for x in np.frombuffer(mv, dtype=it, count=k, offset=pos)]; pos += it.itemsize * k
                rows.append(row)
            if name == "face":
                key = "vertex_indices" if (rows and "vertex_indices" in rows[0]) else "vertex_index"
                polys = [r[key] for r in rows]
            elif name == "tristrips":
                polys = [tri for r in rows for tri in unpack_tstrips(r["vertex_indices"])]
    return _finish_trimesh(verts, polys)


def read_obj(path: str):
    """Wavefront OBJ: `v x y z` and `f a b c ...` (also `t`), 1-based, negative = relative to the vertices read so far;
    only the first integer of an `a/b/c` group counts (TriMesh_io.cu:651-689)."""
    verts, polys = [], []
    with open(path) as f:
        for ln in f:
            t = ln.split()
            if not t or t[0].startswith("#"):
                continue
            if t[0] == "v" and len(t) >= 4:
                verts.append([float(t[1]), float(t[2]), float(t[3])])
            elif t[0] in ("f", "t"):
                face = []
                for g in t[1:]:
                    try:
                        i = int(g.split("/")[0])
                    except ValueError:
                        break
                    face.append(i + len(verts) if i < 0 else i - 1)
                polys.append(face)
    return _finish_trimesh(np.array(verts, dtype=np.float64).reshape(-1, 3), polys)


def _number_tokens(path, skip_first_word=None):
    toks = []
    with open(path) as f:
        for ln in f:
            ln = ln.split("#")[0]
            toks.extend(ln.split())
    if skip_first_word and toks and toks[0] == skip_first_word:
        toks = toks[1:]
    return toks


def read_off(path: str):
    """OFF: `OFF`, `nverts nfaces [nedges]`, the vertices, then `k i0 .. ik-1` per face (TriMesh_io.cu:692-707)."""
    t = _number_tokens(path, "OFF")
    nv, nf = int(t[0]), int(t[1])
    pos = 3
    verts = np.array(t[pos:pos + 3 * nv], dtype=np.float64).reshape(nv, 3)
    pos += 3 * nv
    polys = []
    for _ in range(nf):
        k = int(t[pos]); pos += 1
        polys.append([int(x) for x in t[pos:pos + k]]); pos += k
    return _finish_trimesh(verts, polys)


def read_sm(path: str):
    """Old-style SM: `nverts`, the vertices, `nfaces`, three indices per face (TriMesh_io.cu:710-731)."""
    t = _number_tokens(path)
    nv = int(t[0])
    verts = np.array(t[1:1 + 3 * nv], dtype=np.float64).reshape(nv, 3)
    pos = 1 + 3 * nv
    polys = []
    if pos < len(t):
        nf = int(t[pos]); pos += 1
        polys = [[int(x) for x in t[pos + 3 * i:pos + 3 * i + 3]] for i in range(nf)]
    return _finish_trimesh(verts, polys)


def read_3ds(path: str):
    """3D Studio: little-endian chunks (id u16, length u32).  0x4d4d / 0x3d3d are entered, 0x4000 is entered after its
    zero-terminated name, 0x4100 starts a mesh (indices are relative to its first vertex), 0x4110 = u16 count + float32
    xyz, 0x4120 = u16 count + 4 x u16 per face (a, b, c, flags), everything else is skipped (TriMesh_io.cu:503-576).
    Deviation: upstream's loop `while (!feof(f)) { if (!fread(...)) return false; ...}` returns false at the end of EVERY
    well-formed file (feof is only set by the failing read), so it cannot read 3DS at all; here a clean end is success."""
    import struct
    with open(path, "rb") as f:
        data = f.read()
    pos, mstart, verts, polys = 0, 0, [], []
    while pos + 6 <= len(data):
        cid, clen = struct.unpack_from("<HI", data, pos)
        pos += 6
        if cid in (0x4D4D, 0x3D3D):
            continue
        if cid == 0x4000:
            end = data.find(b"\0", pos)
            if end < 0:
                raise ValueError("truncated 3DS object name")
            pos = end + 1
        elif cid == 0x4100:
            mstart = len(verts)
        elif cid == 0x4110:
            (nv,) = struct.unpack_from("<H", data, pos)
            pos += 2
            verts.extend(np.frombuffer(data, dtype="<f4", count=3 * nv, offset=pos).astype(np.float64).reshape(nv, 3).tolist())
            pos += 12 * nv
        elif cid == 0x4120:
            (nf,) = struct.unpack_from("<H", data, pos)
            pos += 2
            fa = np.frombuffer(data, dtype="<u2", count=4 * nf, offset=pos).reshape(nf, 4)
            polys.extend((fa[:, :3].astype(np.int64) + mstart).tolist())
            pos += 8 * nf
        else:
            pos += max(clen - 6, 0)
    return _finish_trimesh(np.array(verts, dtype=np.float64).reshape(-1, 3), polys)


def read_vvd(path: str):
    """VIVID (Minolta) range scans, big-endian: "VIVID", 127 bytes of header, int32 vertex count, 3 float64 per vertex,
    int32 face count, per face an int32 index count followed by that many int32 indices (TriMesh_io.cu:580-621,
    read_faces_bin with face_len = 4, face_count = 0, face_idx = 4)."""
    import struct
    with open(path, "rb") as f:
        data = f.read()
    if data[:5] != b"VIVID":
        raise ValueError("not a VVD file")
    pos = 5 + 127
    (nv,) = struct.unpack_from(">i", data, pos); pos += 4
    if nv < 0 or pos + 24 * nv > len(data):
        raise ValueError("Couldn't read vertex")
    verts = np.frombuffer(data, dtype=">f8", count=3 * nv, offset=pos).astype(np.float64).reshape(nv, 3)
    pos += 24 * nv
    (nf,) = struct.unpack_from(">i", data, pos); pos += 4
    polys = []
    for _ in range(nf):
        (k,) = struct.unpack_from(">i", data, pos); pos += 4
        polys.append([int(i) for i in struct.unpack_from(">%di" % k, data, pos)]); pos += 4 * k
    return _finish_trimesh(verts, polys)


def read_ray(path: str):
    """Ray-tracer scene text: `#vertex x y z` and `#shape_triangle material a b c` (0-based), every other token ignored
    (TriMesh_io.cu:625-647).  Deviation: upstream scans the coordinates with "%f" into doubles (undefined behaviour: the
    low halves of the doubles get float bit patterns), so its RAY vertices are garbage; here they are parsed as written."""
    with open(path) as f:
        t = f.read().split()
    verts, polys, i = [], [], 0
    while i < len(t):
        if t[i].startswith("#vertex") and i + 3 < len(t):
            verts.append([float(t[i + 1]), float(t[i + 2]), float(t[i + 3])]); i += 4
        elif t[i].startswith("#shape_triangle") and i + 4 < len(t):
            polys.append([int(t[i + 2]), int(t[i + 3]), int(t[i + 4])]); i += 5
        else:
            i += 1
    return _finish_trimesh(np.array(verts, dtype=np.float64).reshape(-1, 3), polys)


def read_trimesh(path: str):
    """Format by the first bytes, as TriMesh::read_helper does (TriMesh_io.cu:160-256)."""
    with open(path, "rb") as f:
        head = f.read(64)
    if not head:
        raise ValueError("Can't read header")
    c = head[:1]
    if head[:3] == b"ply":
        return read_ply(path)
    if head[:2] == b"MM":
        return read_3ds(path)
    if head[:5] == b"VIVID":
        return read_vvd(path)
    if head[:3] == b"OFF":
        return read_off(path)
    if c == b"#":  # the word after '#': material / vertex / shape_... = a ray file, anything else an OBJ comment
        w = head[1:].split()
        if w and (w[0].startswith(b"material") or w[0].startswith(b"vertex") or w[0].startswith(b"shape_")):
            return read_ray(path)
        return read_obj(path)
    if c in (b"v", b"u", b"f", b"g", b"s", b"o"):
        return read_obj(path)
    if c.isdigit():
        return read_sm(path)
    raise ValueError("Unknown file type")


def read_ply_ascii(path: str):
    """Kept name of the first reader (ASCII PLY fixtures of the reference); any PLY flavour is accepted now."""
    return read_ply(path)


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
