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
        else:
            nodes, dx, origin = bb.sdf_grid_layout(c["bounds_min"], c["bounds_max"], c.get("dx", 0.01), c.get("margin", 0.1))
            ix, iy, iz = np.meshgrid(np.arange(nodes[0]), np.arange(nodes[1]), np.arange(nodes[2]), indexing="ij")
            pts = np.stack([origin[0] + dx * ix, origin[1] + dx * iy, origin[2] + dx * iz], axis=-1)
            field = np.ascontiguousarray(np.asarray(c["sdf"](pts.reshape(-1, 3))).reshape(nodes).transpose(2, 1, 0))
            cols.append(O.make_collider("sdf", friction=c.get("friction", 0.0),
                                        sdf=dict(res=nodes, spacing=(dx, dx, dx), origin=origin, field=field)))
    return O.Oracle(sc["spacing"], sc["scale"], sc["domain_min"], sc["domain_max"], cols, **kw)


def sdf_torus(center, R, r):
    """SDF_Torus (src/shapes/sdfs.h:17-22): torus around the y axis."""
    c = np.asarray(center, float)

    def f(p):
        q = p - c
        a = np.sqrt(q[:, 0] ** 2 + q[:, 2] ** 2) - R
        return np.sqrt(a * a + q[:, 1] ** 2) - r
    return f
