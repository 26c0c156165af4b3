"""CPU-only tests of the host-side pieces: data formats, the C-ABI surface, the import contract."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import sci_solver_fem_b200 as fsb
from tests.util import golden, kuhn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
meshio = fsb.meshio


def test_library_loads_and_exports_every_declared_symbol():
    L = fsb.load_library()
    hdr = open(os.path.join(ROOT, "include", "femsolver_b200.h")).read()
    declared = set(re.findall(r"\b(fsb_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(fsb.EXPORTED_SYMBOLS), declared ^ set(fsb.EXPORTED_SYMBOLS)
    raw = C.CDLL(os.path.join(ROOT, "sci-solver_fem_b200", "libfemsolver_b200.so"))
    for name in declared:
        assert hasattr(raw, name), name
    assert L.fsb_version() >= 100


def test_no_cpu_fallback_without_device():
    L = fsb.load_library()
    if L.fsb_device_count() > 0:
        pytest.skip("a device is present")
    with pytest.raises(fsb.FEMSolverError, match="no CPU fallback"):
        fsb.FEMSolver(None)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sci-solver_fem_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "fem_oracle" not in src and "import oracle" not in src and "from oracle" not in src, f


def test_host_quadrature_tables_match_oracle_bitwise():
    from oracle.oracle import lib
    L = fsb.load_library()
    a, b = np.zeros(10), np.zeros(10)
    L.fsb_tet_mass_integrals(a.ctypes.data_as(C.c_void_p))
    lib().orc_tet_mass_integrals(b.ctypes.data_as(C.c_void_p))
    assert np.array_equal(a, b)
    exact = np.array([2, 1, 1, 1, 2, 1, 1, 2, 1, 2]) / 15.0
    assert np.abs(a - exact).max() < 1e-14
    zx, zy, wx, wy = (np.zeros(6) for _ in range(4))
    L.fsb_tri_quadrature(*(z.ctypes.data_as(C.c_void_p) for z in (zx, zy, wx, wy)))
    # Newton stops at |step| < 1e-6 (FEM3D.cu:221), so the rules are Gauss rules only to ~1e-11
    assert abs(wx.sum() - 2.0) < 1e-10 and abs(wy.sum() - 2.0) < 1e-10   # int 1 dx, int (1-y) dy on [-1,1]
    assert zx[0] == -1 and zx[-1] == 1 and zy[0] == -1


def test_mat_roundtrip(tmp_path):
    x = np.random.default_rng(0).normal(size=37)
    p = str(tmp_path / "x.mat")
    meshio.write_mat_array(p, x)
    assert np.array_equal(meshio.read_mat_array(p), x)
    raw = open(p, "rb").read()
    assert raw.startswith(b"MATLAB 5.0 MAT-file, Platform: GLNXA64, Created by SCI-Solver_FEM.")
    assert raw[124:128] == b"\x00\x01IM" and len(raw) == 128 + 8 + 48 + 37 * 8
    import scipy.io
    assert np.array_equal(scipy.io.loadmat(p)["x_h"].ravel(), x)


def test_mat_sparse_fixture_parse():
    g = golden("simple3d")
    ptr, col, val = meshio.csc_to_csr(int(g["A_nrows"]), int(g["A_ncols"]), g["A_jc"], g["A_ir"], g["A_pr"])
    assert ptr[-1] == 5000 and np.array_equal(col, np.arange(5000)) and np.allclose(val, 8 * np.pi ** 2 + 1, rtol=1e-6)


def test_node_ele_roundtrip_and_float_rounding(tmp_path):
    v, t = kuhn(3, h=0.1)
    base = str(tmp_path / "m")
    meshio.write_node_ele(base, v, t, labels=np.arange(len(t)) % 7)
    v2, t2, lab2 = meshio.read_node_ele(base)
    assert np.array_equal(t2, t) and np.array_equal(lab2, np.arange(len(t)) % 7)
    assert np.array_equal(v2, v.astype(np.float32).astype(np.float64))   # %f into float (tetmesh.cu:285-292)
    # one-based files are shifted down
    with open(base + ".ele") as f:
        lines = f.read().splitlines()
    with open(base + ".ele", "w") as f:
        f.write(lines[0] + "\n")
        for ln in lines[1:]:
            p = ln.split()
            f.write(" ".join([p[0]] + [str(int(q) + 1) for q in p[1:5]] + p[5:]) + "\n")
    _, t3, _ = meshio.read_node_ele(base)
    assert np.array_equal(t3, t)


def test_ply_roundtrip(tmp_path):
    v, f = meshio.grid_tri(4, 3)
    p = str(tmp_path / "g.ply")
    meshio.write_ply_ascii(p, v, f)
    v2, f2 = meshio.read_ply_ascii(p)
    assert np.array_equal(v2, v) and np.array_equal(f2, f)
    g = golden("simple2d")
    assert g["verts"].shape == (2500, 3) and g["tris"].shape == (4802, 3)


def test_kuhn_cube_counts_and_orientation():
    v, t = kuhn(6)
    assert v.shape == (343, 3) and t.shape == (6 * 216, 4)
    X = v[t]
    vol = np.linalg.det(X[:, 1:] - X[:, :1]) / 6
    assert np.all(vol > 0) and abs(vol.sum() - 1) < 1e-6
    lab = meshio.kuhn_cell_labels(16, block=8)
    assert lab.shape == (6 * 16 ** 3,) and set(np.unique(lab)) <= set(range(1, 7))


def test_upstream_callers_compile_against_the_dropin_headers(tmp_path):
    """Source compatibility of the drop-in layer: the reference's own example and gtest sources are compiled
    (syntax + types only, in place, nothing copied) against dropin/FEMSolver.h and the gtest shim.  Needs the
    read-only reference checkout, so it is skipped on the GPU box."""
    import shutil
    import subprocess
    ref = "/root/reference/src"
    if not os.path.isdir(ref) or not shutil.which("g++"):
        pytest.skip("reference checkout or g++ not available")
    dropin = os.path.join(ROOT, "sci-solver_fem_b200", "dropin")
    sources = [os.path.join(ref, "test", f) for f in ("sanity2D.cc", "sanity3D.cc", "tetVol.cc")]
    sources += [os.path.join(ref, "examples", f) for f in ("example1.cu", "example2.cu")]
    for src in sources:
        cmd = ["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-x", "c++", "-I", dropin, "-DTEST_DATA_DIR=fsb_test_data_dir()", src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, src + "\n" + r.stderr[-2000:]


def test_bench_cpu_baseline_helper():
    """bench.py's CPU arm: the oracle run returns the stage times, the iteration count and, on request, the
    same solve on one host thread (what the bench line reports as cpu_baseline / cpu_baseline.single_thread)."""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    r = bench.run_oracle(10, 1, 0, single_thread_solve=True)
    assert r["n"] == 11 ** 3 and r["iters"] > 0 and r["relres"] <= 1e-8
    assert r["t_solve"] > 0 and r["t_solve_1thread"] > 0 and r["threads"] >= 1
    assert bench.run_oracle(10, 1, 0)["t_solve_1thread"] is None
    assert bench.algorithmic_bytes("pre_smooth", 10, 100, nnz_in=60) == 12 * 60 + 4 * 11 + 240
    assert bench.algorithmic_bytes("spmv_dot", 10, 100) == 12 * 100 + 4 * 11 + 160


def _mesh_with_polygons():
    """Small surface with triangles, two quads (one per diagonal choice) and a pentagon."""
    verts = np.array([[0, 0, 0], [1, 0, 0], [0.8, 0.8, 0], [0, 1, 0], [2, 0, 0.25], [2.5, 1, 0], [1, 2, 0], [0, 2, 0.5], [-1, 1, 0], [3, 2.5, 1]],
                     dtype=np.float64)
    polys = [[0, 1, 2, 3], [1, 4, 5, 2], [3, 2, 6], [3, 6, 7, 8, 0], [4, 9, 5]]
    return verts, polys


def _write_mesh_files(d, verts, polys):
    """The same mesh as PLY (ascii / binary LE / binary BE with extra properties and a leading element), OBJ, OFF, SM."""
    import struct
    files = {}
    nv, nf = len(verts), len(polys)
    hdr = ("ply\nformat %s 1.0\ncomment test\nelement extra 2\nproperty int a\nelement vertex %d\nproperty float x\nproperty float y\n"
           "property float z\nproperty uchar red\nproperty double w\nelement face %d\nproperty list uchar int vertex_indices\n"
           "property float q\nend_header\n")
    with open(os.path.join(d, "a.ply"), "w") as f:
        f.write(hdr % ("ascii", nv, nf))
        f.write("7\n9\n")
        for v in verts:
            f.write("%.9g %.9g %.9g 200 1.5\n" % tuple(v))
        for p in polys:
            f.write("%d %s 0.5\n" % (len(p), " ".join(map(str, p))))
    files["ply_ascii"] = os.path.join(d, "a.ply")
    for tag, e, name in (("le", "<", "binary_little_endian"), ("be", ">", "binary_big_endian")):
        path = os.path.join(d, "b_%s.ply" % tag)
        with open(path, "wb") as f:
            f.write((hdr % (name, nv, nf)).encode())
            f.write(struct.pack(e + "ii", 7, 9))
            for v in verts:
                f.write(struct.pack(e + "fffBd", v[0], v[1], v[2], 200, 1.5))
            for p in polys:
                f.write(struct.pack(e + "B%dif" % len(p), len(p), *p, 0.5))
        files["ply_" + tag] = path
    with open(os.path.join(d, "m.obj"), "w") as f:
        f.write("# comment\n")
        for v in verts:
            f.write("v %.9g %.9g %.9g\n" % tuple(v))
        for k, p in enumerate(polys):
            if k % 2 == 0:
                f.write("f " + " ".join("%d/%d/%d" % (i + 1, i + 1, i + 1) for i in p) + "\n")
            else:
                f.write("f " + " ".join(str(i - nv) for i in p) + "\n")     # negative = relative to the end
    files["obj"] = os.path.join(d, "m.obj")
    with open(os.path.join(d, "m.off"), "w") as f:
        f.write("OFF\n# comment\n%d %d 0\n" % (nv, nf))
        for v in verts:
            f.write("%.9g %.9g %.9g\n" % tuple(v))
        for p in polys:
            f.write("%d %s\n" % (len(p), " ".join(map(str, p))))
    files["off"] = os.path.join(d, "m.off")
    return files


def test_triangle_mesh_formats_python_and_dropin_agree(tmp_path):
    """TriMesh::read's format detection and polygon tessellation (TriMesh_io.cu:160-256, 1239-1270) for PLY (ascii and
    both binary byte orders), OBJ, OFF and SM: the Python readers and the drop-in's C++ readers must return the same
    mesh, and that mesh must be the hand-tessellated one."""
    import subprocess
    verts, polys = _mesh_with_polygons()
    v32 = verts.astype(np.float32).astype(np.float64)
    expect = []
    for p in polys:
        expect += meshio.tessellate(verts, p)
    # the two quads exercise both diagonal choices, the pentagon the fan
    assert expect[:2] == [(0, 1, 2), (0, 2, 3)] and expect[2:4] == [(4, 5, 2), (4, 2, 1)] and len(expect) == 2 + 2 + 1 + 3 + 1
    files = _write_mesh_files(str(tmp_path), verts, polys)
    tri_only = [list(t) for t in expect]
    with open(os.path.join(str(tmp_path), "m.sm"), "w") as f:     # SM has no polygons: write the triangles
        f.write("%d\n" % len(verts))
        for v in verts:
            f.write("%.9g %.9g %.9g\n" % tuple(v))
        f.write("%d\n" % len(tri_only))
        for t in tri_only:
            f.write("%d %d %d\n" % tuple(t))
    files["sm"] = os.path.join(str(tmp_path), "m.sm")
    exe = os.path.join(ROOT, "sci-solver_fem_b200", "dropin", "bin", "mesh_info")
    have_exe = os.path.exists(exe)
    for kind, path in files.items():
        v, t = meshio.read_trimesh(path)
        assert np.allclose(v, v32 if kind.startswith("ply_") and kind != "ply_ascii" else verts, rtol=0, atol=1e-6), kind
        assert [tuple(x) for x in t.tolist()] == expect, kind
        if have_exe:
            r = subprocess.run([exe, path], capture_output=True, text=True)
            assert r.returncode == 0, kind + r.stderr
            out = r.stdout.split("\n")
            assert int(out[0]) == len(v) and int(out[1]) == len(t), kind
            cs = sum((i % 97 + 1) * (j + 1) * v[i, j] for i in range(len(v)) for j in range(3))
            fs = sum((i % 89 + 1) * (j + 1) * int(t[i, j]) for i in range(len(t)) for j in range(3))
            assert abs(float(out[2]) - cs) <= 1e-9 * max(1.0, abs(cs)), kind
            assert int(out[3]) == fs, kind
    with open(os.path.join(str(tmp_path), "bad.xyz"), "w") as f:
        f.write("?? nothing\n")
    with pytest.raises(ValueError):
        meshio.read_trimesh(os.path.join(str(tmp_path), "bad.xyz"))
    if have_exe:
        assert subprocess.run([exe, os.path.join(str(tmp_path), "bad.xyz")], capture_output=True).returncode == 1



def _cmp_with_dropin(path, v, t):
    """The C++ reader of the drop-in (bin/mesh_info) must see the same mesh as the Python reader."""
    import subprocess
    exe = os.path.join(ROOT, "sci-solver_fem_b200", "dropin", "bin", "mesh_info")
    if not os.path.exists(exe):
        return
    r = subprocess.run([exe, path], capture_output=True, text=True)
    assert r.returncode == 0, path + r.stderr
    out = r.stdout.split("\n")
    assert int(out[0]) == len(v) and int(out[1]) == len(t), path
    cs = sum((i % 97 + 1) * (j + 1) * v[i, j] for i in range(len(v)) for j in range(3))
    fs = sum((i % 89 + 1) * (j + 1) * int(t[i, j]) for i in range(len(t)) for j in range(3))
    assert abs(float(out[2]) - cs) <= 1e-9 * max(1.0, abs(cs)), path
    assert int(out[3]) == fs, path


def test_remaining_triangle_mesh_formats(tmp_path):
    """The rest of TriMesh::read's formats (TriMesh_io.cu:146-256): PLY triangle strips (ascii and binary), 3D Studio,
    VIVID (big-endian) and ray-tracer scenes, in Python and in the drop-in's C++ reader; a PLY range grid gives vertices
    without faces, as upstream (its grid reader is commented out)."""
    import struct
    verts, polys = _mesh_with_polygons()
    v32 = verts.astype(np.float32).astype(np.float64)
    nv = len(verts)
    d = str(tmp_path)
    # --- PLY tristrips: two strips, the second stitched with a degenerate triangle
    strips = [0, 1, 2, 3, 4, -1, 5, 6, 7, 7, 8, -1]
    want = [(0, 1, 2), (2, 1, 3), (2, 3, 4), (5, 6, 7)]   # (6,7,7) and (7,7,8) are degenerate: dropped
    assert [tuple(x) for x in meshio.unpack_tstrips(strips)] == want
    head = "ply\nformat %s 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\nelement tristrips 1\nproperty list int int vertex_indices\nend_header\n"
    pa = os.path.join(d, "strips_ascii.ply")
    with open(pa, "w") as f:
        f.write(head % ("ascii", nv))
        for v in verts:
            f.write("%.9g %.9g %.9g\n" % tuple(v))
        f.write("%d %s\n" % (len(strips), " ".join(map(str, strips))))
    for order, e in (("binary_little_endian", "<"), ("binary_big_endian", ">")):
        pb = os.path.join(d, "strips_%s.ply" % order)
        with open(pb, "wb") as f:
            f.write((head % (order, nv)).encode())
            f.write(verts.astype(e + "f4").tobytes())
            f.write(struct.pack(e + "i", len(strips)) + np.array(strips, dtype=e + "i4").tobytes())
        v, t = meshio.read_trimesh(pb)
        assert np.allclose(v, v32, atol=1e-7) and [tuple(x) for x in t.tolist()] == want
        _cmp_with_dropin(pb, v, t)
    v, t = meshio.read_trimesh(pa)
    assert np.allclose(v, verts, atol=1e-6) and [tuple(x) for x in t.tolist()] == want
    _cmp_with_dropin(pa, v, t)
    # --- PLY range grid: vertices only
    pg = os.path.join(d, "grid.ply")
    with open(pg, "w") as f:
        f.write("ply\nformat ascii 1.0\nobj_info num_cols 3\nobj_info num_rows 3\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                "element range_grid 9\nproperty list uchar int vertex_indices\nend_header\n" % nv)
        for v in verts:
            f.write("%.9g %.9g %.9g\n" % tuple(v))
        for k in range(9):
            f.write("1 %d\n" % k)
    v, t = meshio.read_trimesh(pg)
    assert len(v) == nv and len(t) == 0
    _cmp_with_dropin(pg, v, t)
    # --- 3D Studio: main > model > two objects, the second with its own vertex numbering and a sub-chunk to skip
    tris = [t for p in polys for t in meshio.tessellate(verts, p)]
    ha, hb = nv // 2, nv - nv // 2
    ta = [t for t in tris if max(t) < ha]
    tb = [tuple(i - ha for i in t) for t in tris if min(t) >= ha]
    def chunk(cid, payload):
        return struct.pack("<HI", cid, 6 + len(payload)) + payload
    def mesh(vs, ts, extra=b""):
        vc = chunk(0x4110, struct.pack("<H", len(vs)) + np.asarray(vs, dtype="<f4").tobytes())
        fc = chunk(0x4120, struct.pack("<H", len(ts)) + b"".join(struct.pack("<4H", a, b, c, 7) for a, b, c in ts) + extra)
        return chunk(0x4100, vc + fc)
    objs = chunk(0x4000, b"first\0" + mesh(verts[:ha], ta)) + chunk(0xAFFF, b"material junk") + \
        chunk(0x4000, b"second\0" + mesh(verts[ha:], tb, chunk(0x4150, struct.pack("<%dI" % max(len(tb), 1), *([1] * max(len(tb), 1))))))
    p3 = os.path.join(d, "m.3ds")
    with open(p3, "wb") as f:
        f.write(chunk(0x4D4D, chunk(0x0002, struct.pack("<I", 3)) + chunk(0x3D3D, objs)))
    v, t = meshio.read_trimesh(p3)
    assert np.allclose(v, v32, atol=1e-7)
    assert [tuple(x) for x in t.tolist()] == ta + [tuple(i + ha for i in q) for q in tb]
    _cmp_with_dropin(p3, v, t)
    # --- VIVID: big-endian doubles, polygons with their index counts
    pv = os.path.join(d, "m.vvd")
    with open(pv, "wb") as f:
        f.write(b"VIVID" + bytes(127) + struct.pack(">i", nv) + verts.astype(">f8").tobytes() + struct.pack(">i", len(polys)))
        for p in polys:
            f.write(struct.pack(">i", len(p)) + np.array(p, dtype=">i4").tobytes())
    v, t = meshio.read_trimesh(pv)
    assert np.array_equal(v, verts) and [tuple(x) for x in t.tolist()] == tris
    _cmp_with_dropin(pv, v, t)
    # --- RAY: '#vertex x y z' / '#shape_triangle material a b c' among other tokens
    pr = os.path.join(d, "m.ray")
    with open(pr, "w") as f:
        f.write("#material 0 0 0  1 1 1\n#camera 0 0 5\n")
        for v in verts:
            f.write("#vertex %.17g %.17g %.17g\n" % tuple(v))
        for a, b, c in tris:
            f.write("#shape_triangle 0 %d %d %d\n" % (a, b, c))
    v, t = meshio.read_trimesh(pr)
    assert np.array_equal(v, verts) and [tuple(x) for x in t.tolist()] == tris
    _cmp_with_dropin(pr, v, t)
    # an OBJ that starts with an ordinary comment is still an OBJ
    po = os.path.join(d, "c.obj")
    with open(po, "w") as f:
        f.write("# exported by something\nv 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    v, t = meshio.read_trimesh(po)
    assert len(v) == 3 and t.tolist() == [[0, 1, 2]]
    _cmp_with_dropin(po, v, t)


def test_bench_parity_block_and_byte_models(monkeypatch):
    """bench.py's parity block (oracle golden + single-GPU comparison of a sharded solve) accepts rounding-level
    differences, rejects a wrong iteration count / history / solution, and the stage byte models add up."""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, 1000)
    hist = np.geomspace(1.0, 1e-9, 30)
    g = {"iterations": 29, "resid_history": list(hist), "sample_idx": [0, 10, 999], "x_samples": [float(x[i]) for i in (0, 10, 999)],
         "x_norm2": float(np.linalg.norm(x)), "levels": [1000, 40]}
    monkeypatch.setattr(bench, "golden_for", lambda N: (g, "tests/golden/fake.json"))
    levels = [(1000, 15000), (40, 900)]
    ok = bench.parity_block(7, x * (1 + 1e-13), 29, hist * (1 + 1e-9), levels, x_single=x, iters_single=29)
    assert ok["ok"] and ok["checks"] == ["oracle_golden", "vs_single_gpu"] and ok["oracle_golden"]["iters_equal"]
    assert not bench.parity_block(7, x, 33, hist, levels)["ok"]                      # iterations off by more than 2
    assert not bench.parity_block(7, x, 29, hist * (1 + 1e-4), levels)["ok"]         # residual history differs
    assert not bench.parity_block(7, x * (1 + 1e-4), 29, hist, levels)["ok"]         # solution differs
    assert not bench.parity_block(7, x, 29, hist, [(1000, 15000), (41, 900)])["ok"]  # different aggregates
    assert not bench.parity_block(7, x, 29, hist, levels, x_single=x + 1e-3, iters_single=29)["ok"]
    monkeypatch.setattr(bench, "golden_for", lambda N: (None, None))
    assert bench.parity_block(7, x, 29, hist, levels)["oracle_golden"] is None
    assert bench.setup_bytes(levels, [3000, 0], [4000, 0]) == 24 * 15000 + 10 * 11000 + 12 * 4000 + 24 * 3000 + 12 * (15000 + 6000 + 900) + 16 * 1000
    assert bench.host_threads() >= 1
    assert bench.workload_name(255).startswith("Kuhn tet cube N=255 (16777216 DOF, 99488250 tets)")


def test_oracle_goldens_are_complete():
    """The committed full-size goldens (tests/golden/make_oracle_golden.py) carry what the GPU tests and bench.py read, and the
    problem builder shared with them is self-consistent."""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("make_oracle_golden", os.path.join(ROOT, "tests", "golden", "make_oracle_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    for N, variant, iters in ((118, "", 33), (255, "", 39), (149, "contrast", 36), (149, "anisotropic", 60)):
        g = json.load(open(m.golden_path(N, variant)))
        assert g["cube"] == N and g["iterations"] == iters and g["levels"][0] == (N + 1) ** 3
        assert len(g["resid_history"]) >= iters and len(g["sample_idx"]) == len(g["x_samples"]) >= 3
        assert g["x_norm2"] > 0 and g["b_norm2"] > 0 and max(g["sample_idx"]) == (N + 1) ** 3 - 1
        if variant != "anisotropic":
            assert g["relres"] <= 1e-8
    v, t, lab, xs, prm = m.problem(6, "contrast")
    assert v.shape == (343, 3) and t.shape == (6 * 216, 4) and lab.shape == (6 * 216,) and xs.shape == (343,) and prm["maxIters"] == 400
    v, t, lab, xs, prm = m.problem(6, "anisotropic")
    assert lab is None and abs(v[:, 2].max() - 1.0 / 64.0) < 1e-7 and prm["tolerance"] == 1e-30
