"""Scenes shared by the tests, __graft_entry__.smoke() and bench.py (synthetic, reproducible).

Constants follow the reference's scene scripts: test_pcisph3_dam_break
(src/tests/test_pcisph_extra.cpp:1102-1169: spacing 0.02, kernel scale 1.8, reversed container box,
fluid box with initial velocity) and the survey's probe scene.
"""
import numpy as np

import bubbles_b200 as bb
from bubbles_b200 import emitter


def f32(a):
    """FP32-representable values held in float64 (identical input for the FP64 oracle and the FP32 engine)."""
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def block_scene(container, fluid_size, fluid_center, v0, spacing=0.02, scale=1.8, jitter=0.001, seed=1, dt=7e-4,
                extra_colliders=()):
    b = emitter.ParticleSetBuilder3()
    lo = np.asarray(fluid_center, float) - np.asarray(fluid_size, float) / 2
    hi = np.asarray(fluid_center, float) + np.asarray(fluid_size, float) / 2
    em = emitter.VolumeParticleEmitter3(emitter.box_inside(fluid_center, fluid_size), lo, hi, spacing, v0, jitter, seed)
    em.Emit(b)
    half = np.asarray(container, float) / 2
    return dict(spacing=spacing, scale=scale, dt=dt, domain_min=-half, domain_max=half,
                colliders=[dict(kind="box", size=tuple(container), reverse=True, friction=0.0)] + list(extra_colliders),
                pos=f32(b.positions), vel=f32(b.velocities))


def probe_scene():
    """0.6^3 reversed box, 0.2 x 0.3 x 0.2 fluid block at (0.1, -0.1, 0.1), v0 = (0, -1, 0): ~3k particles."""
    return block_scene((0.6, 0.6, 0.6), (0.2, 0.3, 0.2), (0.1, -0.1, 0.1), (0, -1, 0))


def dam_break_scene(n_target=1.0e6, spacing=0.02, scale=1.8, jitter=0.0, seed=1):
    """test_pcisph3_dam_break scaled so that the BCC count is ~n_target (N = 2 V / s^3)."""
    # reference proportions: container 3.25 x 3.0 x 3.25, fluid 1.25 x 2.25 x 1.25 (domainScaling 2.5)
    v_ref = 1.25 * 2.25 * 1.25
    k = (n_target * spacing ** 3 / 2.0 / v_ref) ** (1.0 / 3.0)
    container = np.array([3.25, 3.0, 3.25]) * k
    fluid = np.array([1.25, 2.25, 1.25]) * k
    xof = (container[0] - fluid[0]) / 2 - spacing
    zof = (container[2] - fluid[2]) / 2 - spacing
    yof = (container[1] - fluid[1]) / 2 - spacing
    return block_scene(container, fluid, (xof, -yof, zof), (0, -6, 0), spacing, scale, jitter, seed, dt=7.2e-4)


def dam_break_geometry(n_target=1.0e6, spacing=0.02):
    """Container / fluid box of test_pcisph3_dam_break scaled to ~n_target BCC particles."""
    v_ref = 1.25 * 2.25 * 1.25
    k = (n_target * spacing ** 3 / 2.0 / v_ref) ** (1.0 / 3.0)
    container = np.array([3.25, 3.0, 3.25]) * k
    fluid = np.array([1.25, 2.25, 1.25]) * k
    center = np.array([(container[0] - fluid[0]) / 2 - spacing, -((container[1] - fluid[1]) / 2 - spacing),
                       (container[2] - fluid[2]) / 2 - spacing])
    return container, fluid, center


def dam_break_scene_slab(n_target, rank, world, spacing=0.02, scale=1.8, obstacle=None):
    """The jitter-free dam-break block WITHOUT materialising all of it on every rank: the BCC lattice is built
    layer by layer (z), the per-plane histogram comes from the layer sizes, and a rank only generates the
    layers whose cell plane it owns.  Returns the scene dict with pos / vel / ids of this rank's share, the
    global count `n_global`, the grid, the z plan and the plane histogram.  Same points, same ids (lattice
    order) as dam_break_scene(jitter=0)."""
    container, fluid, center = dam_break_geometry(n_target, spacing)
    half = container / 2
    colliders = [dict(kind="box", size=tuple(container), reverse=True, friction=0.0)]
    if obstacle is not None:
        colliders.append(obstacle(container, fluid, center))
    sc = dict(spacing=spacing, scale=scale, dt=7.2e-4, domain_min=-half, domain_max=half, colliders=colliders)
    grid = bb.UtilBuildGridForDomain(sc["domain_min"], sc["domain_max"], spacing, scale)
    lo, hi = center - fluid / 2, center + fluid / 2
    ext = np.abs(hi - lo)
    hs = spacing / 2
    layers = []  # (z, x values, y values, first id, count, cell plane); the keep tests are emitter.box_inside per axis
    k, shifted, first = 0, False, 0
    hist = np.zeros(grid.n[2], dtype=np.int64)
    gmin, glen = grid.min[2], grid.cell_len[2]
    while k * hs <= ext[2]:
        off = hs if shifted else 0.0
        z = k * hs + lo[2]
        ny = int(np.floor((ext[1] - off) / spacing + 1e-9)) + 2
        nx = int(np.floor((ext[0] - off) / spacing + 1e-9)) + 2
        x = np.arange(nx) * spacing + off
        y = np.arange(ny) * spacing + off
        x, y = x[x <= ext[0]] + lo[0], y[y <= ext[1]] + lo[1]
        zf = float(np.float32(z))
        xk = x[np.abs(x - center[0]) <= fluid[0] / 2]
        yk = y[np.abs(y - center[1]) <= fluid[1] / 2]
        cnt = len(xk) * len(yk) if abs(z - center[2]) <= fluid[2] / 2 else 0
        plane = int(np.floor((zf - gmin) / glen))
        plane = min(max(plane, 0), grid.n[2] - 1)
        layers.append((z, xk, yk, first, cnt, plane))
        hist[plane] += cnt
        first += cnt
        shifted = not shifted
        k += 1
    zb = bb.plan_slabs(hist, world)
    pos, ids = [], []
    for z, xk, yk, f0, cnt, plane in layers:
        if cnt and zb[rank] <= plane < zb[rank + 1]:
            yy, xx = np.meshgrid(yk, xk, indexing="ij")
            pos.append(np.stack([xx, yy, np.full_like(xx, z)], axis=-1).reshape(-1, 3).astype(np.float32))
            ids.append(np.arange(f0, f0 + cnt, dtype=np.int32))
    sc["pos"] = np.concatenate(pos) if pos else np.zeros((0, 3), np.float32)
    sc["ids"] = np.concatenate(ids) if ids else np.zeros(0, np.int32)
    sc["vel"] = np.tile(np.array([0, -6, 0], dtype=np.float32), (len(sc["pos"]), 1))
    sc.update(n_global=int(first), grid=grid, z_bounds=zb, hist=hist)
    return sc


def _xf(c):
    t = c.get("translate")
    return bb.Translate(*t) if t is not None else None


def engine_colliders(sc):
    out = []
    for c in sc["colliders"]:
        if c["kind"] == "box":
            s = bb.MakeBox(_xf(c), c["size"], c.get("reverse", False))
        elif c["kind"] == "sphere":
            s = bb.MakeSphere(_xf(c), c["radius"], c.get("reverse", False))
        elif c["kind"] == "mesh":
            s = bb.MakeMesh(c["vertices"], c["triangles"], c["sdf"], c.get("reverse", False))
        else:
            s = bb.MakeSDFShape(c["bounds_min"], c["bounds_max"], c["sdf"], c.get("dx", 0.01), c.get("margin", 0.1))
        s.friction = c.get("friction", 0.0)
        out.append(s)
    return out


def make_engine(sc, max_particles=None, **kw):
    grid = bb.UtilBuildGridForDomain(sc["domain_min"], sc["domain_max"], sc["spacing"], sc["scale"])
    eng = bb.Engine(grid, sc["spacing"], sc["scale"], max_particles or max(1, len(sc["pos"])), **kw)
    eng.set_colliders(engine_colliders(sc))
    return eng


def make_oracle(sc, **kw):
    from oracle import oracle as O
    cols = []
    for c in sc["colliders"]:
        t = c.get("translate")
        m = O.translate(*t) if t is not None else O.IDENTITY
        if c["kind"] == "box":
            cols.append(O.make_collider("box", m, size=c["size"], reverse=c.get("reverse", False), friction=c.get("friction", 0.0)))
        elif c["kind"] == "sphere":
            cols.append(O.make_collider("sphere", m, radius=c["radius"], reverse=c.get("reverse", False), friction=c.get("friction", 0.0)))
        elif c["kind"] == "mesh":
            cols.append(O.make_collider("mesh", friction=c.get("friction", 0.0), reverse=c.get("reverse", False),
                                        mesh=(c["vertices"], c["triangles"]), sdf=c["sdf"]))
        else:
            nodes, dx, origin = bb.sdf_grid_layout(c["bounds_min"], c["bounds_max"], c.get("dx", 0.01), c.get("margin", 0.1))
            ix, iy, iz = np.meshgrid(np.arange(nodes[0]), np.arange(nodes[1]), np.arange(nodes[2]), indexing="ij")
            pts = np.stack([origin[0] + dx * ix, origin[1] + dx * iy, origin[2] + dx * iz], axis=-1)
            field = np.ascontiguousarray(np.asarray(c["sdf"](pts.reshape(-1, 3))).reshape(nodes).transpose(2, 1, 0))
            cols.append(O.make_collider("sdf", friction=c.get("friction", 0.0),
                                        sdf=dict(res=nodes, spacing=(dx, dx, dx), origin=origin, field=field)))
    return O.Oracle(sc["spacing"], sc["scale"], sc["domain_min"], sc["domain_max"], cols, **kw)


def torus_mesh(center, R, r, nu=32, nv=16):
    """A closed, consistently oriented triangle mesh of a torus around the y axis (stand-in for the absent whale / dragon
    .obj files): (vertices [nu * nv, 3] float64, triangles [2 * nu * nv, 3] int32)."""
    u = np.arange(nu) * 2 * np.pi / nu
    v = np.arange(nv) * 2 * np.pi / nv
    U, V = np.meshgrid(u, v, indexing="ij")
    P = np.stack([(R + r * np.cos(V)) * np.cos(U), r * np.sin(V), (R + r * np.cos(V)) * np.sin(U)], -1).reshape(-1, 3) + np.asarray(center, float)
    idx = lambda i, j: (i % nu) * nv + (j % nv)
    T = []
    for i in range(nu):
        for j in range(nv):
            a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
            T += [(a, b, c), (a, c, d)]
    return np.ascontiguousarray(P, dtype=np.float64), np.array(T, dtype=np.int32)


def mesh_collider_from_golden(g, friction=0.1):
    """scene collider entry of a mesh collider whose SDF grid was generated by the REFERENCE (tests/golden/mesh_collider.npz)"""
    return dict(kind="mesh", vertices=g["vertices"], triangles=g["triangles"], friction=friction,
                sdf=dict(res=tuple(int(x) for x in g["sdf_res"]), spacing=tuple(g["sdf_meta"][:3]), origin=tuple(g["sdf_meta"][3:]), field=g["sdf_field"]))


def torus_obstacle(container, fluid, center):
    """Config 3 stand-in for the absent whale / dragon meshes: an analytic SDF torus
    (SDF_Torus, src/shapes/sdfs.h:17-22) baked with MakeSDFShape at dx = 0.01, lying on the floor right beside
    the fluid block so that the collapsing front runs into it."""
    R, r = 0.30 * fluid[0], 0.08 * fluid[0]
    c = np.array([center[0] - fluid[0] / 2 - R - r - 0.05, -container[1] / 2 + r + 0.02, center[2]])
    return dict(kind="sdf", bounds_min=tuple(c - np.array([R + r, r, R + r])), bounds_max=tuple(c + np.array([R + r, r, R + r])),
                sdf=sdf_torus(c, R, r), dx=0.01, margin=0.1, friction=0.0)


def sdf_torus(center, R, r):
    """SDF_Torus (src/shapes/sdfs.h:17-22): torus around the y axis."""
    c = np.asarray(center, float)

    def f(p):
        q = p - c
        a = np.sqrt(q[:, 0] ** 2 + q[:, 2] ** 2) - R
        return np.sqrt(a * a + q[:, 1] ** 2) - r
    return f
