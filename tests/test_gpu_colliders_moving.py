"""SURVEY.md 8(f)4, cheap half: colliders that MOVE or are TOGGLED between sub-steps on the GPU.

bbx_update_collider = Shape::Update + Shape::SetVelocities (src/core/shape.cpp:290-305: new ObjectToWorld /
WorldToObject, linear and angular velocity used by VelocityAt in the sphere's closest-point query, sphere.cpp:52-70);
bbx_set_collider_active = ColliderSet3::SetActive (src/core/collider.cpp:232-236).  The oracle side of both is pinned
bit-exactly against the unmodified reference (test_moving_sphere_and_toggled_collider_response_bit_exact); here the
engine follows the oracle through a scripted motion, state resynced before every traced sub-step (identical inputs),
integer results bit-exact, fields within the tolerances of test_gpu_parity.py.  Also on slab engines, with the
collider moving across the cut.
"""
import numpy as np
import pytest

import bubbles_b200 as bb
import parity_gate as pg
import scenes
from oracle import oracle as O

pytestmark = pytest.mark.gpu

RADIUS = 0.07


def _scene():
    # block falling onto the floor of a 0.6^3 box; a sphere obstacle (index 1) and a box obstacle (index 2) under it
    extra = [dict(kind="sphere", radius=RADIUS, translate=(0.1, -0.25, 0.1), friction=0.3),
             dict(kind="box", size=(0.08, 0.05, 0.08), translate=(0.02, -0.27, 0.16), friction=0.1)]
    return scenes.block_scene((0.6, 0.6, 0.6), (0.2, 0.3, 0.2), (0.1, -0.1, 0.1), (0, -2, 0), extra_colliders=extra)


def _sphere_at(center, linvel, angvel, active=True):
    """(engine collider, oracle collider) of the obstacle sphere at `center` with the given velocities"""
    e = bb.MakeSphere(bb.Translate(*center), RADIUS)
    e.friction = 0.3
    e.linear_velocity[:] = linvel
    e.angular_velocity[:] = angvel
    e.active = int(active)
    o = O.make_collider("sphere", O.translate(*center), radius=RADIUS, friction=0.3, linvel=linvel, angvel=angvel, active=active)
    return e, o


def _set_oracle_collider(orc, index, c):
    import ctypes as C
    C.memmove(C.byref(orc._colliders[index]), C.byref(c), C.sizeof(O.Collider))


def _path(k, z0=0.1):
    """scripted motion: the sphere sweeps in +z / -x while spinning (a keyframed TransformSequence sampled per sub-step)"""
    t = 7e-4 * k
    lin = (-0.8, 0.1 * np.cos(40 * t), 2.0)
    center = (0.1 + lin[0] * t, -0.25 + 0.0025 * np.sin(40 * t), z0 + lin[2] * t)
    return tuple(float(np.float64(c)) for c in center), tuple(float(v) for v in lin), (0.0, 25.0, -10.0)


def test_moving_spinning_sphere_and_toggled_box_follow_the_oracle():
    sc = _scene()
    eng = scenes.make_engine(sc)
    orc = scenes.make_oracle(sc)
    eng.set_particles(sc["pos"], sc["vel"])
    orc.set_particles(sc["pos"], sc["vel"])
    dt = sc["dt"]
    ext = float(np.max(sc["domain_max"] - sc["domain_min"]))
    hits_sphere = hits_box_on = 0
    k = 0
    for block in range(6):
        for _ in range(12):                      # the collider moves EVERY sub-step on both sides
            c, lin, ang = _path(k)
            ec, oc = _sphere_at(c, lin, ang)
            eng.update_collider(1, ec)
            _set_oracle_collider(orc, 1, oc)
            orc.substep_pcisph(dt)
            eng.step_pcisph(dt)
            k += 1
        box_on = block % 2 == 0                  # ColliderSet3::SetActive between frames
        eng.set_collider_active(2, box_on)
        orc._colliders[2].active = int(box_on)
        c, lin, ang = _path(k)
        ec, oc = _sphere_at(c, lin, ang)
        eng.update_collider(1, ec)
        _set_oracle_collider(orc, 1, oc)
        pg.sync_engine_from_oracle(eng, orc)
        pos_in = orc.a["pos"].copy()
        r, tr = pg.gate_substep(eng, orc, dt, ext)
        k += 1
        assert r["ok"], (block, r)
        moved = np.abs(tr["pos_out"] - (pos_in + dt * tr["vel_out"])).max(axis=1) > 1e-9
        d_sphere = np.linalg.norm(tr["pos_out"] - np.asarray(c), axis=1) - RADIUS
        hits_sphere += int((moved & (d_sphere < 2.5 * sc["spacing"])).sum())
        q = np.abs(tr["pos_out"] - np.array([0.02, -0.27, 0.16])) - np.array([0.04, 0.025, 0.04])
        near_box = (q.max(axis=1) < 2.5 * sc["spacing"])
        if box_on:
            hits_box_on += int((moved & near_box).sum())
    assert hits_sphere > 0, "the moving sphere never touched the fluid"
    assert hits_box_on > 0, "the box obstacle never touched the fluid while active"
    # the carried velocity (Shape::VelocityAt) matters: the sphere's tangential spin shows up in the response
    assert eng.stats().nan_count == 0
    eng.close()


def test_update_collider_rejects_type_change_and_bad_index():
    sc = _scene()
    eng = scenes.make_engine(sc)
    with pytest.raises(bb.BbxError):
        eng.update_collider(1, bb.MakeBox(None, (0.1, 0.1, 0.1)))
    with pytest.raises(bb.BbxError):
        eng.update_collider(7, bb.MakeSphere(None, 0.1))
    with pytest.raises(bb.BbxError):
        eng.set_collider_active(-1, 1)
    eng.close()


def test_collider_moving_across_a_slab_cut_matches_single_domain_bit_for_bit():
    """Every slab engine gets the same update; the sphere travels in +z through the cuts of a 3-slab group.
    (It starts just outside the block: a particle deep inside a collider at t = 0 is ejected by up to the sphere's
    radius = 2 cell planes in one sub-step, more than the one-plane halo of a slab engine can follow -- that case is
    BBX_ERR_OUT_OF_DOMAIN by design, checked in the last test of this file.)"""
    Z0 = -0.075
    sc = _scene()
    sc["colliders"][1]["translate"] = (0.1, -0.25, Z0)
    grid = bb.UtilBuildGridForDomain(sc["domain_min"], sc["domain_max"], sc["spacing"], sc["scale"])
    hist = bb.plane_histogram(grid, sc["pos"])
    zb = bb.plan_slabs(hist, 3)
    n = len(sc["pos"])
    grp = bb.LocalSlabGroup(grid, sc["spacing"], sc["scale"], zb, n, ghost_capacity=n)
    grp.set_colliders(scenes.engine_colliders(sc))
    grp.set_particles(sc["pos"], sc["vel"])
    one = scenes.make_engine(sc)
    one.set_particles(sc["pos"], sc["vel"])
    dt = sc["dt"]
    lo = grid.min[2] + zb[1] * grid.cell_len[2]
    crossed = False
    z_prev = None
    for k in range(90):
        c, lin, ang = _path(k, Z0)
        ec, _ = _sphere_at(c, lin, ang)
        one.update_collider(1, ec)
        for e in grp.engines:
            e.update_collider(1, ec)
        if k == 45:
            one.set_collider_active(2, 0)
            for e in grp.engines:
                e.set_collider_active(2, 0)
        one.step_pcisph(dt)
        grp.step_pcisph(dt)
        if z_prev is not None and (z_prev - RADIUS < lo) != (c[2] + RADIUS < lo):
            crossed = True
        z_prev = c[2]
    assert crossed or (Z0 - RADIUS < lo < _path(89, Z0)[0][2] + RADIUS), "the sphere never reached a slab cut"
    for f in (bb.POSITION, bb.VELOCITY, bb.DENSITY):
        assert np.array_equal(grp.download(f, np.float32), one.download(f, np.float32)), f"field {f}"
    cc, co = grp.export_cells()
    c1, o1 = one.export_cells()
    assert np.array_equal(cc, c1) and np.array_equal(co, o1)
    grp.close()
    one.close()


def test_two_plane_jump_on_a_slab_engine_is_reported_not_silently_dropped():
    """A particle that moves two cell planes in one sub-step (here: ejected from deep inside a collider) cannot be
    followed by the one-plane halo of a slab engine.  The single-domain engine takes the reference's route (jump
    detection -> full rebuild); the slab engine must return BBX_ERR_OUT_OF_DOMAIN from the stepping API instead of
    stepping on without the particle (device-side errors never reached the caller)."""
    sc = _scene()                                   # the block starts overlapping the sphere obstacle
    grid = bb.UtilBuildGridForDomain(sc["domain_min"], sc["domain_max"], sc["spacing"], sc["scale"])
    zb = bb.plan_slabs(bb.plane_histogram(grid, sc["pos"]), 3)
    n = len(sc["pos"])
    grp = bb.LocalSlabGroup(grid, sc["spacing"], sc["scale"], zb, n, ghost_capacity=n)
    grp.set_colliders(scenes.engine_colliders(sc))
    grp.set_particles(sc["pos"], sc["vel"])
    with pytest.raises(bb.BbxError) as ei:
        for _ in range(4):
            grp.step_pcisph(sc["dt"])
    assert ei.value.code == bb.ERR_OUT_OF_DOMAIN
    # a fresh particle set clears the sticky error
    calm = scenes.block_scene((0.6, 0.6, 0.6), (0.2, 0.3, 0.2), (0.0, 0.1, 0.0), (0, -1, 0))
    grp.set_particles(calm["pos"], calm["vel"])
    for _ in range(3):
        grp.step_pcisph(sc["dt"])
    assert all(s.nan_count == 0 for s in grp.stats())
    grp.close()
