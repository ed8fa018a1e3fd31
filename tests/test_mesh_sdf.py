"""Mesh colliders, host side: the facade's MakeMesh + GenerateShapeSDF (bubbles_b200/host/mesh_sdf.h) against
the grid the unmodified reference bakes for the same closed mesh (tests/golden/mesh_collider.npz: GenerateShapeSDF ->
SetNodeSDFKernel: BVH closest distance, sign by ray parity) -- node layout, bounds and every field value bit for bit -- and
MeshClosestDistance against the reference's Shape::ClosestDistance at 4 000 points.  Host code; no GPU involved."""
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
TOOL = os.path.join(ROOT, "bubbles_b200", "lib", "frame_tool")


def bake(vertices, triangles, dx, margin, queries=None):
    wd = tempfile.mkdtemp(prefix="bake_")
    mesh, out = os.path.join(wd, "mesh.bin"), os.path.join(wd, "out.bin")
    with open(mesh, "wb") as f:
        f.write(np.array([len(vertices), len(triangles)], dtype=np.int64).tobytes())
        f.write(np.ascontiguousarray(vertices, dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(triangles, dtype=np.int32).tobytes())
    cmd = [TOOL, "--bake", mesh, out, repr(float(dx)), repr(float(margin))]
    if queries is not None:
        q = os.path.join(wd, "q.bin")
        with open(q, "wb") as f:
            f.write(np.array([len(queries)], dtype=np.int64).tobytes()); f.write(np.ascontiguousarray(queries, dtype=np.float64).tobytes())
        cmd.append(q)
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    res = np.frombuffer(raw, dtype=np.int64, count=3)
    meta = np.frombuffer(raw, dtype=np.float64, count=10, offset=24)
    n = int(res.prod())
    field = np.frombuffer(raw, dtype=np.float64, count=n, offset=24 + 80)
    dist = np.frombuffer(raw, dtype=np.float64, offset=24 + 80 + 8 * n) if queries is not None else None
    return res, meta, field, dist


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def test_baked_grid_is_bit_identical_to_the_reference_bake():
    g = np.load(os.path.join(G, "mesh_collider.npz"))
    res, meta, field, dist = bake(g["vertices"], g["triangles"], 0.01, 0.1, g["q_pos"])
    assert np.array_equal(res, g["sdf_res"])
    assert np.array_equal(bits(meta[0:1]), bits(g["sdf_meta"][0:1])) and np.array_equal(bits(meta[1:4]), bits(g["sdf_meta"][3:6]))   # spacing, origin
    assert np.array_equal(bits(meta[4:10]), bits(g["sdf_bounds"]))                                                                    # Shape::GetBounds
    assert (np.sign(field) == np.sign(g["sdf_field"])).all(), "inside / outside differs at %d nodes" % (np.sign(field) != np.sign(g["sdf_field"])).sum()
    assert np.array_equal(bits(field), bits(g["sdf_field"]))
    assert (g["sdf_field"] < 0).sum() > 100 and (g["sdf_field"] > 0).sum() > 100          # the grid has an inside and an outside
    assert np.array_equal(bits(dist), bits(g["distance"]))                                    # Shape::ClosestDistance at 4 000 points


def test_inside_test_is_direction_independent_on_a_closed_mesh():
    """An octahedron: nodes on its symmetry planes send their ray through edges and vertices -- the grazing rule must turn the
    ray instead of miscounting; the signs must match the analytic |x| + |y| + |z| < r."""
    r = 0.25
    v = np.array([[r, 0, 0], [-r, 0, 0], [0, r, 0], [0, -r, 0], [0, 0, r], [0, 0, -r]], dtype=np.float64)
    t = np.array([[0, 2, 4], [2, 1, 4], [1, 3, 4], [3, 0, 4], [2, 0, 5], [1, 2, 5], [3, 1, 5], [0, 3, 5]], dtype=np.int32)
    res, meta, field, _ = bake(v, t, 0.05, 0.2)
    dx, org = meta[0], meta[1:4]
    k, j, i = np.meshgrid(np.arange(res[2]), np.arange(res[1]), np.arange(res[0]), indexing="ij")
    p = np.stack([org[0] + dx * i, org[1] + dx * j, org[2] + dx * k], axis=-1).reshape(-1, 3)
    l1 = np.abs(p).sum(1)
    clear = np.abs(l1 - r) > 1e-9                              # nodes exactly on the surface can go either way
    assert ((field < 0) == (l1 < r))[clear].all()
    exact = np.abs(l1 - r) / np.sqrt(3.0)                      # distance to the face plane: exact wherever the foot point is on the face
    inner = l1 < r
    assert np.abs(np.abs(field[inner]) - np.maximum(exact[inner], 1e-5)).max() < 1e-12


def test_obj_loader_matches_the_reference_loader():
    """LoadObj (bubbles_b200/host/obj_loader.h) against the reference's loader (src/third/obj_loader.cpp) through the harness: a
    file with comments, CRLF line ends, every corner syntax (i, i/j, i//k, i/j/k), negative indices, quads, unused vertices,
    numbers in plain / exponent form -- vertices (first-use order, bit for bit) and triangle indices must agree."""
    import pytest
    from oracle import oracle as O
    if not O.ref_available():
        pytest.skip("oracle/_ref/bbref not built")
    g = np.load(os.path.join(G, "mesh_collider.npz"))
    v, t = g["vertices"], g["triangles"]
    wd = tempfile.mkdtemp(prefix="obj_")
    path = os.path.join(wd, "mesh.obj")
    rng = np.random.default_rng(11)
    lines = ["# generated by tests/test_mesh_sdf.py", "mtllib none.mtl", "o blob", "v 9.5 9.5 9.5", "v -1.25e-3 4.0E+1 .5"]   # two vertices no face uses
    fmts = ["%.17g", "%.9f", "%.6e", "%g"]
    for k, p in enumerate(v):
        lines.append("v " + " ".join(fmts[k % 4] % x for x in p) + ("\r" if k % 7 == 0 else ""))
    lines += ["vn 0 0 1", "vn 0 1 0", "vt 0.5 0.5", "vt 0.25 0.75", "usemtl skin"]
    nv = len(v) + 2
    for k, tri in enumerate(t[:600]):
        a = tri + 3                                   # 1-based, behind the two unused vertices
        style = k % 5
        if style == 0: c = ["%d" % i for i in a]
        elif style == 1: c = ["%d/%d" % (i, 1 + k % 2) for i in a]
        elif style == 2: c = ["%d//%d" % (i, 1 + k % 2) for i in a]
        elif style == 3: c = ["%d/%d/%d" % (i, 1 + k % 2, 2 - k % 2) for i in a]
        else: c = ["%d" % (i - nv - 1) for i in a]      # negative: relative to the vertices read so far
        lines.append("f " + " ".join(c) + ("  " if k % 3 == 0 else ""))
    for k in range(40):                                # quads
        q = rng.integers(3, nv + 1, 4)
        lines.append("f " + " ".join("%d" % i for i in q))
    open(path, "w", newline="").write("\n".join(lines) + "\n")
    O.run_ref([f"load_obj {path} {wd}/r_"], wd)
    out = os.path.join(wd, "out.bin")
    r = subprocess.run([TOOL, "--obj", path, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    n, m = np.frombuffer(raw, dtype=np.int64, count=2)
    pts = np.frombuffer(raw, dtype=np.float64, count=3 * n, offset=16).reshape(n, 3)
    tri = np.frombuffer(raw, dtype=np.int32, count=3 * m, offset=16 + 24 * n).reshape(m, 3)
    rp, rt = np.load(os.path.join(wd, "r_points.npy")), np.load(os.path.join(wd, "r_triangles.npy"))
    assert (n, m) == (len(rp), len(rt)) and m == 600 + 80
    assert np.array_equal(bits(pts), bits(rp)) and np.array_equal(tri, rt)
    assert n <= len(v)                                   # the two unused vertices are gone
