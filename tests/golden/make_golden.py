"""Regenerates tests/golden/*.npz from the reference's own test fixtures.

Run HERE (the container that mounts /root/reference); the GPU box has no /root/reference,
so the compact .npz files are what travels.  Content: the meshes exactly as the reference
parses them (float-rounded coordinates for .node, %lf for PLY), and the .mat matrices /
vectors of the three gtests (sanity2D.cc, sanity3D.cc, tetVol.cc), plus the reference's
CubeMesh_size256step16 pair.  No reference SOURCE code is copied — only test DATA.

    python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src/test/test_data"

spec = importlib.util.spec_from_file_location("meshio", os.path.join(ROOT, "sci-solver_fem_b200", "meshio.py"))
meshio = importlib.util.module_from_spec(spec)
spec.loader.exec_module(meshio)


def mat_sparse(name):
    nr, nc, jc, ir, pr = meshio.read_mat_sparse(os.path.join(REF, name))
    return dict(nrows=nr, ncols=nc, jc=jc, ir=ir, pr=pr)


def main():
    out = {}
    # tetVol: the one genuine known-answer fixture (tetVol.cc)
    v, t, lab = meshio.read_node_ele(os.path.join(REF, "tetVol"))
    A = mat_sparse("tetVolA.mat")
    np.savez_compressed(os.path.join(HERE, "tetVol.npz"), verts=v, tets=t, labels=lab,
                        b=meshio.read_mat_array(os.path.join(REF, "tetVolb.mat")),
                        ans=meshio.read_mat_array(os.path.join(REF, "tetVolAns.mat")),
                        **{"A_" + k: np.asarray(x) for k, x in A.items()})
    # simple (sanity3D.cc)
    v, t, lab = meshio.read_node_ele(os.path.join(REF, "simple"))
    A = mat_sparse("simple.mat")
    np.savez_compressed(os.path.join(HERE, "simple3d.npz"), verts=v, tets=t, labels=lab,
                        b=meshio.read_mat_array(os.path.join(REF, "simpleb.mat")),
                        ans=meshio.read_mat_array(os.path.join(REF, "simpleAns.mat")),
                        **{"A_" + k: np.asarray(x) for k, x in A.items()})
    # simple.ply (sanity2D.cc)
    v, f = meshio.read_ply_ascii(os.path.join(REF, "simple.ply"))
    A = mat_sparse("simpleTri.mat")
    np.savez_compressed(os.path.join(HERE, "simple2d.npz"), verts=v, tris=f,
                        b=meshio.read_mat_array(os.path.join(REF, "simpleTrib.mat")),
                        ans=meshio.read_mat_array(os.path.join(REF, "simpleTriAns.mat")),
                        **{"A_" + k: np.asarray(x) for k, x in A.items()})
    # Example1 default input and its conforming twin
    for nm in ("CubeMesh_size256step16", "CubeMesh_size256step16_correct"):
        v, t, lab = meshio.read_node_ele(os.path.join(REF, nm))
        np.savez_compressed(os.path.join(HERE, nm + ".npz"), verts=v, tets=t, labels=lab)
    v, f = meshio.read_ply_ascii(os.path.join(REF, "sphere_290verts.ply"))
    np.savez_compressed(os.path.join(HERE, "sphere_290verts.npz"), verts=v, tris=f)
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith(".npz"):
            print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == "__main__":
    sys.exit(main())
