"""GPU parity tests: the CUDA engine (through the C ABI, bubbles_b200/lib/libbbx.so) against the
oracle (oracle/bbx_oracle.c, pinned bit-exact to the unmodified reference by
tests/test_oracle_vs_reference.py) on identical, FP32-representable inputs.

Bars: integer / index results (cell ids, per-cell chain order, neighbour lists, flags) BIT-EXACT;
floating-point fields within the FP32-engine-vs-FP64-reference tolerances written below.
"""
import numpy as np
import pytest

import bubbles_b200 as bb
import scenes

pytestmark = pytest.mark.gpu

# tolerances (engine FP32 vs oracle FP64, same inputs, one sub-step); measured errors are ~10x smaller
TOL_RHO = 2e-5        # relative to rho0
TOL_FORCE = 2e-4      # relative to max |f| of the field
TOL_POS = 1e-6        # absolute, in units of the domain extent
TOL_VEL = 2e-5        # relative to max |v|
TOL_PRESSURE = 2e-3   # relative to max p (p = delta * (rho* - rho0) amplifies the FP32 density rounding by rho0/err)


def relmax(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _pair(sc, **kw):
    eng = scenes.make_engine(sc, **kw)
    orc = scenes.make_oracle(sc)
    eng.set_particles(sc["pos"], sc["vel"])
    orc.set_particles(sc["pos"], sc["vel"])
    return eng, orc


def _compare_phase_by_phase(eng, orc, dt, sc):
    """Run one sub-step phase by phase on both sides; returns the oracle trace."""
    rho0 = 1000.0
    tr = orc.trace_pcisph(dt)
    eng.run_phase(bb.PHASE_GRID, dt)
    cc, co = eng.export_cells()
    assert np.array_equal(cc, tr["cell_count"]), "cell counts differ"
    assert np.array_equal(co, tr["cell_order"]), "per-cell chain order differs"
    eng.run_phase(bb.PHASE_DENSITY, dt)
    cnt, ids = eng.export_neighbors()
    assert np.array_equal(cnt, tr["nbr_count"]), "neighbour counts differ"
    assert np.array_equal(ids, tr["nbr_ids"]), "neighbour lists differ"
    assert np.array_equal(eng.download(bb.NEIGHBOR_COUNT), tr["nbr_count"])
    rho = eng.download(bb.DENSITY)
    assert np.abs(rho - tr["density"]).max() / rho0 < TOL_RHO
    eng.run_phase(bb.PHASE_FORCE_NP, dt)
    assert relmax(eng.download(bb.FORCE_NP), tr["force_np"]) < TOL_FORCE
    ext = float(np.max(sc["domain_max"] - sc["domain_min"]))
    assert np.abs(eng.download(bb.PRED_POSITION) - tr["pos_pred"]).max() / ext < TOL_POS
    eng.run_phase(bb.PHASE_PRESSURE, dt)
    assert np.abs(eng.download(bb.PRED_DENSITY) - tr["density_pred"]).max() / rho0 < TOL_RHO
    pmax = max(tr["pressure"].max(), 1e-30)
    assert np.abs(eng.download(bb.PRESSURE) - tr["pressure"]).max() / pmax < TOL_PRESSURE
    assert np.abs(eng.download(bb.DENSITY_ERROR) - tr["density_error"]).max() / rho0 < TOL_RHO
    eng.run_phase(bb.PHASE_PRESSURE_FORCE, dt)
    fscale = max(np.abs(tr["force_p"]).max(), np.abs(tr["force_np"]).max())
    assert np.abs(eng.download(bb.PRESSURE_FORCE) - tr["force_p"]).max() / fscale < 10 * TOL_FORCE
    eng.run_phase(bb.PHASE_INTEGRATE, dt)
    assert np.abs(eng.download(bb.POSITION) - tr["pos_out"]).max() / ext < TOL_POS
    vmax = np.abs(tr["vel_out"]).max()
    assert np.abs(eng.download(bb.VELOCITY) - tr["vel_out"]).max() / vmax < 10 * TOL_VEL
    st = eng.stats()
    assert st.rebuild_flag == tr["rebuild_flag_out"]
    assert st.neighbor_overflow == tr["overflow"]
    return tr


def test_setup_scalars_match_oracle():
    sc = scenes.probe_scene()
    eng, orc = _pair(sc)
    assert eng.mass == orc.P.mass                       # ComputeMass, FP64 host arithmetic: bit-exact
    assert eng.delta(7e-4) == orc.delta(7e-4)           # ComputeDelta
    assert abs(eng.mass - 0.0040391407688202411) < 1e-18  # value printed by the unmodified reference
    assert abs(eng.delta(7e-4) - 602.06523574685036) < 1e-9


def test_initial_full_rebuild_is_ascending_id():
    sc = scenes.probe_scene()
    eng, orc = _pair(sc)
    cc, co = eng.export_cells()
    assert np.array_equal(cc, orc.arr("cell_count"))
    assert np.array_equal(co, orc.arr("cell_order"))
    # shuffled ids: chains must still be ascending id inside each cell
    rng = np.random.default_rng(3)
    perm = rng.permutation(len(sc["pos"]))
    eng.set_particles(sc["pos"][perm], sc["vel"][perm])
    orc.set_particles(sc["pos"][perm], sc["vel"][perm])
    cc, co = eng.export_cells()
    assert np.array_equal(cc, orc.arr("cell_count"))
    assert np.array_equal(co, orc.arr("cell_order"))


def test_first_substep_phase_by_phase():
    sc = scenes.probe_scene()
    eng, orc = _pair(sc)
    _compare_phase_by_phase(eng, orc, sc["dt"], sc)


def test_history_dependent_order_from_injected_state():
    """Chain order after many incremental updates, neighbour lists and all fields, at sub-steps 10, 20, 30:
    both sides restart from the oracle's state rounded to FP32 (identical inputs), including the chains."""
    sc = scenes.probe_scene()
    eng, orc = _pair(sc)
    dt = sc["dt"]
    for block in range(3):
        for _ in range(10):
            orc.substep_pcisph(dt)
        pos, vel = scenes.f32(orc.a["pos"]), scenes.f32(orc.a["vel"])
        orc.a["pos"][:] = pos
        orc.a["vel"][:] = vel
        eng.overwrite_state(pos, vel)
        eng.inject_chains(orc.arr("cell_count"), orc.arr("cell_order"))
        eng.set_rebuild_flag(orc.S.rebuild_flag)
        _compare_phase_by_phase(eng, orc, dt, sc)


def test_free_running_cells_and_lists_stay_exact_when_resynced_every_step():
    """40 sub-steps; after each one the engine state is replaced by the oracle's (rounded to FP32 on both
    sides) so inputs stay identical while the chains evolve independently on each side."""
    sc = scenes.probe_scene()
    eng, orc = _pair(sc)
    dt = sc["dt"]
    for step in range(40):
        orc.substep_pcisph(dt)
        eng.step_pcisph(dt)
        cc, co = eng.export_cells()
        assert np.array_equal(cc, orc.arr("cell_count")), f"step {step}"
        assert np.array_equal(co, orc.arr("cell_order")), f"step {step}"
        cnt, ids = eng.export_neighbors()
        assert np.array_equal(cnt, orc.a["nbr_count"]), f"step {step}"
        assert np.array_equal(ids, orc.a["nbr_ids"]), f"step {step}"
        pos, vel = scenes.f32(orc.a["pos"]), scenes.f32(orc.a["vel"])
        orc.a["pos"][:] = pos
        orc.a["vel"][:] = vel
        eng.overwrite_state(pos, vel)


def test_trajectory_tolerance_100_substeps():
    """Free-running 100 sub-steps (includes the block hitting the floor): |dx| <= 1e-2 * spacing for >= 99.9 %
    of the particles (SURVEY.md Appendix D.4 gate)."""
    sc = scenes.probe_scene()
    eng, orc = _pair(sc)
    for _ in range(100):
        orc.substep_pcisph(sc["dt"])
    eng.step_many(sc["dt"], 100)
    d = np.linalg.norm(eng.download(bb.POSITION) - orc.a["pos"], axis=1)
    assert np.quantile(d, 0.999) <= 1e-2 * sc["spacing"], np.quantile(d, [0.5, 0.99, 0.999, 1.0])
    st = eng.stats()
    assert st.nan_count == 0 and st.substeps == 100


def test_big_move_flag_forces_full_rebuild():
    sc = scenes.probe_scene()
    sc["vel"] = scenes.f32(np.tile([0.0, -60.0, 0.0], (len(sc["pos"]), 1)))  # 0.042 per step > 0.9 * 0.036
    eng, orc = _pair(sc)
    dt = sc["dt"]
    orc.substep_pcisph(dt)
    eng.step_pcisph(dt)
    assert orc.S.rebuild_flag == 1 and eng.stats().rebuild_flag == 1
    pos, vel = scenes.f32(orc.a["pos"]), scenes.f32(orc.a["vel"])
    orc.a["pos"][:] = pos
    orc.a["vel"][:] = vel
    eng.overwrite_state(pos, vel)
    orc.substep_pcisph(dt)
    eng.step_pcisph(dt)
    assert eng.stats().full_rebuild == 1
    cc, co = eng.export_cells()
    assert np.array_equal(cc, orc.arr("cell_count")) and np.array_equal(co, orc.arr("cell_order"))


def test_neighbor_cap_100_is_mirrored():
    """A compressed block (spacing 0.55 x nominal) gives > 100 candidates inside h: the stored lists must be
    the first 100 in the reference's traversal order, and the density the sum over exactly those."""
    sc = scenes.block_scene((0.6, 0.6, 0.6), (0.12, 0.12, 0.12), (0.0, 0.0, 0.0), (0, 0, 0), jitter=0.0)
    center = sc["pos"].mean(axis=0)
    sc["pos"] = scenes.f32(center + (sc["pos"] - center) * 0.55)
    eng, orc = _pair(sc)
    tr = orc.trace_pcisph(sc["dt"])
    assert tr["overflow"] > 0
    eng.run_phase(bb.PHASE_GRID, sc["dt"])
    eng.run_phase(bb.PHASE_DENSITY, sc["dt"])
    cnt, ids = eng.export_neighbors()
    assert np.array_equal(cnt, tr["nbr_count"]) and cnt.max() == 100
    assert np.array_equal(ids, tr["nbr_ids"])
    assert np.abs(eng.download(bb.DENSITY) - tr["density"]).max() / tr["density"].max() < TOL_RHO
    assert eng.stats().neighbor_overflow == tr["overflow"]
    # (such neighbourhoods exceed the 512-candidate stage of the list build: the chunked path is what ran)
    assert eng.stats().max_candidates > 512


@pytest.mark.parametrize("kind", ["sphere_container", "sphere_obstacle", "box_obstacle", "sdf_torus", "mesh_torus"])
def test_colliders_phase_by_phase(kind):
    extra, container = [], (0.6, 0.6, 0.6)
    if kind == "mesh_torus":   # triangle mesh + the SDF grid the REFERENCE generated for it (tests/golden/mesh_collider.npz)
        import os
        extra = [scenes.mesh_collider_from_golden(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mesh_collider.npz")), friction=0.1)]
    if kind == "sphere_obstacle":
        extra = [dict(kind="sphere", radius=0.08, translate=(0.1, -0.27, 0.1), friction=0.2)]
    elif kind == "box_obstacle":
        extra = [dict(kind="box", size=(0.1, 0.06, 0.1), translate=(0.1, -0.27, 0.1), friction=0.1)]
    elif kind == "sdf_torus":
        extra = [dict(kind="sdf", bounds_min=(-0.05, -0.30, -0.05), bounds_max=(0.25, -0.22, 0.25),
                      sdf=scenes.sdf_torus((0.1, -0.26, 0.1), 0.09, 0.03), dx=0.01, margin=0.1, friction=0.0)]
    sc = scenes.block_scene(container, (0.2, 0.3, 0.2), (0.1, -0.1, 0.1), (0, -2, 0), extra_colliders=extra)
    if kind == "sphere_container":
        sc["colliders"] = [dict(kind="sphere", radius=0.29, reverse=True, friction=0.0)]
    eng, orc = _pair(sc)
    dt = sc["dt"]
    hits = 0
    for block in range(4):
        for _ in range(15):
            orc.substep_pcisph(dt)
        pos, vel = scenes.f32(orc.a["pos"]), scenes.f32(orc.a["vel"])
        orc.a["pos"][:] = pos
        orc.a["vel"][:] = vel
        eng.overwrite_state(pos, vel)
        eng.inject_chains(orc.arr("cell_count"), orc.arr("cell_order"))
        eng.set_rebuild_flag(orc.S.rebuild_flag)
        tr = _compare_phase_by_phase(eng, orc, dt, sc)
        free = tr["pos_out"] - (pos + dt * tr["vel_out"])
        hits += int((np.abs(free).max(axis=1) > 1e-9).sum())
    assert hits > 0, "scene never touched the collider"


def test_mesh_collider_distance_through_the_device_bvh():
    """Shape::MeshClosestDistance on the device (BVH built by bbx_set_colliders, nearest-child-first traversal) against the
    reference's values at 4 000 points (golden), and against the brute-force oracle at points far outside the mesh where
    the traversal prunes hardest.  FP64 on both sides; the device may contract a * b + c into FMAs: 1e-12 relative."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mesh_collider.npz"))
    sc = scenes.probe_scene()
    sc["colliders"] = sc["colliders"] + [scenes.mesh_collider_from_golden(g)]
    eng, orc = _pair(sc)
    d = eng.collider_distance(1, g["q_pos"])
    assert np.abs(d - g["distance"]).max() <= 1e-12 * max(1.0, np.abs(g["distance"]).max())
    rng = np.random.default_rng(4)
    far = rng.uniform(-0.29, 0.29, size=(2000, 3))
    assert np.abs(eng.collider_distance(1, far) - orc.mesh_closest_distance(1, far)).max() <= 1e-12
    # the other families through the same entry (signed distances)
    assert np.abs(eng.collider_distance(0, far) + (0.3 - np.abs(far).max(axis=1))).max() < 1e-12   # reversed container box
    with pytest.raises(bb.BbxError):
        eng.collider_distance(5, far)
    eng.close()


def test_correct_mode_iterates_and_reduces_density_error():
    sc = scenes.block_scene((0.6, 0.6, 0.6), (0.2, 0.3, 0.2), (0.1, -0.15, 0.1), (0, -3, 0))
    eng = scenes.make_engine(sc, reference_compat=False)
    orc = scenes.make_oracle(sc)
    eng.set_particles(sc["pos"], sc["vel"])
    orc.set_particles(sc["pos"], sc["vel"])
    its_e, its_o = [], []
    for _ in range(30):
        its_o.append(orc.substep_pcisph(sc["dt"], compat=False))
        eng.step_pcisph(sc["dt"])
        its_e.append(eng.stats().pcisph_iterations)
    assert max(its_o) > 1, "scene too calm to need a second iteration"
    assert its_e == its_o
    d = np.linalg.norm(eng.download(bb.POSITION) - orc.a["pos"], axis=1)
    assert np.quantile(d, 0.999) <= 1e-2 * sc["spacing"]


def test_sph_step_matches_jacobi_oracle():
    sc = scenes.probe_scene()
    eng, orc = _pair(sc)
    dt = 1.44e-4
    for _ in range(20):
        orc.substep_sph(dt)
        eng.step_sph(dt)
    ext = float(np.max(sc["domain_max"] - sc["domain_min"]))
    assert np.abs(eng.download(bb.POSITION) - orc.a["pos"]).max() / ext < 20 * TOL_POS
    assert relmax(eng.download(bb.VELOCITY), orc.a["vel"]) < 1e-3
    assert np.abs(eng.download(bb.DENSITY) - orc.a["density"]).max() / 1000.0 < 10 * TOL_RHO


def test_sph_step_against_the_reference_run_golden():
    """a20 pinned: tests/golden/sph_run.npz is the UNMODIFIED reference's SphSolver3 on one thread (Gauss-Seidel, the only
    deterministic mode of its CPU path; the oracle's restatement of it is bit-exact, test_oracle_vs_reference.py).  The
    engine runs the race-free Jacobi form: density and EOS pressure of the first sub-step within the FP32 tolerance of the
    reference's own values, trajectories within the stated Jacobi-vs-Gauss-Seidel bound, Chamfer distance
    as in the reference's resources/chamfer.py."""
    import os
    from scipy.spatial import KDTree
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sph_run.npz"))
    sc = scenes.probe_scene()
    sc["colliders"] = sc["colliders"] + [dict(kind="sphere", radius=0.08, translate=(0.1, -0.27, 0.1), friction=0.2)]
    sc["pos"], sc["vel"] = scenes.f32(g["s0_pos"]), scenes.f32(g["s0_vel"])
    eng = scenes.make_engine(sc)
    eng.set_particles(sc["pos"], sc["vel"])
    dt, s = 1.44e-4, sc["spacing"]
    eng.step_sph(dt)
    # (the positions were rounded to FP32 on the way in: 6e-8 relative, amplified by the stiff Tait EOS in the pressure)
    assert np.abs(eng.download(bb.DENSITY) - g["s1_density"]).max() / 1000.0 < 10 * TOL_RHO
    pmax = np.abs(g["s1_pressure"]).max()
    assert np.abs(eng.download(bb.PRESSURE) - g["s1_pressure"]).max() / pmax < 5e-3
    assert np.array_equal(eng.download(bb.NEIGHBOR_COUNT), g["s1_nbr_count"])
    eng.step_many(dt, 19, bb.SOLVER_SPH)
    d = np.linalg.norm(eng.download(bb.POSITION) - g["s20_pos"], axis=1)
    assert d.max() <= 0.03 * s, d.max() / s
    eng.step_many(dt, 100, bb.SOLVER_SPH)
    p = eng.download(bb.POSITION)
    d = np.linalg.norm(p - g["s120_pos"], axis=1)
    assert d.max() <= 0.25 * s, d.max() / s
    ch = KDTree(p).query(g["s120_pos"])[0].mean() + KDTree(g["s120_pos"]).query(p)[0].mean()
    assert ch <= 0.1 * s, ch / s
    assert eng.stats().nan_count == 0
    eng.close()


def test_advance_cfl_substep_count_matches_oracle():
    sc = scenes.probe_scene()
    eng, orc = _pair(sc)
    frame = 1.0 / 240.0
    total = 0
    for _ in range(2):
        remaining, count = frame, 0
        while remaining > np.float64(np.float32(0.0001)):
            nsteps = orc.number_of_time_steps(remaining, 5.0)
            dt = remaining / nsteps
            orc.substep_pcisph(dt)
            remaining -= dt
            count += 1
        sub, ms = eng.advance(frame)
        assert sub == count
        total += sub
    d = np.linalg.norm(eng.download(bb.POSITION) - orc.a["pos"], axis=1)
    assert np.quantile(d, 0.999) <= 1e-2 * sc["spacing"]
    assert eng.stats().substeps == total


def test_determinism_two_runs_bit_identical():
    sc = scenes.probe_scene()
    outs = []
    for _ in range(2):
        eng = scenes.make_engine(sc)
        eng.set_particles(sc["pos"], sc["vel"])
        eng.step_many(sc["dt"], 25)
        outs.append((eng.download(bb.POSITION, np.float32), eng.download(bb.VELOCITY, np.float32)))
        eng.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_empty_and_tiny_inputs():
    sc = scenes.probe_scene()
    eng = scenes.make_engine(sc, max_particles=64)
    eng.set_particles(np.zeros((0, 3)), np.zeros((0, 3)))
    eng.step_pcisph(sc["dt"])
    assert eng.n == 0
    one = scenes.f32([[0.0, 0.0, 0.0]])
    eng.set_particles(one, np.zeros((1, 3)))
    orc = scenes.make_oracle(sc)
    orc.set_particles(one, np.zeros((1, 3)))
    for _ in range(3):
        eng.step_pcisph(sc["dt"])
        orc.substep_pcisph(sc["dt"])
    assert np.abs(eng.download(bb.POSITION) - orc.a["pos"]).max() < 1e-7
    cnt, ids = eng.export_neighbors()
    assert cnt[0] == 1 and ids[0, 0] == 0
    with pytest.raises(bb.BbxError):
        eng.set_particles(np.zeros((65, 3)), np.zeros((65, 3)))


def test_domain_corner_and_face_particles_hash_like_reference():
    """Particles exactly on the domain faces / corners exercise ExtremeEpsilon (grid.h:259-270)."""
    sc = scenes.probe_scene()
    g = bb.UtilBuildGridForDomain(sc["domain_min"], sc["domain_max"], sc["spacing"], sc["scale"])
    lo, hi = np.array(g.min[:]), np.array(g.max[:])
    inner_lo, inner_hi = lo + 0.05, hi - 0.05
    rng = np.random.default_rng(5)
    pts = rng.uniform(inner_lo, inner_hi, size=(256, 3))
    # snap some coordinates onto cell boundaries (multiples of the cell length from the grid min)
    k = rng.integers(2, 15, size=(256, 3))
    snap = lo + k * np.array(g.cell_len[:])
    mask = rng.random((256, 3)) < 0.5
    pts = np.where(mask, snap, pts)
    pts[0] = lo          # f32(lo) lies within 1e-8 of the face: ExtremeEpsilon applies
    pts[1] = hi
    pts[2] = (lo[0], hi[1], 0.0)
    pts = scenes.f32(pts)
    pts = np.minimum(np.maximum(pts, lo), hi)  # FP32 rounding may step outside by < 1 ulp
    pts = scenes.f32(pts)
    sc2 = dict(sc, pos=pts, vel=np.zeros_like(pts), colliders=sc["colliders"])
    eng, orc = _pair(sc2)
    cc, co = eng.export_cells()
    assert np.array_equal(cc, orc.arr("cell_count")) and np.array_equal(co, orc.arr("cell_order"))


def test_append_particles_between_steps_matches_reference_chains():
    """bbx_append_particles after stepping = ContinuousParticleSetBuilder3::AddParticle + Commit (grid.h:1409-1441):
    the new ids join the tail of their cells' chains, the old chains stay.  The oracle side of this is pinned against
    the unmodified reference (tests/golden/append_run.npz); here the engine follows the oracle through two appends
    with the state resynced every step (identical inputs), chains and lists bit-exact."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "append_run.npz"))
    sc = scenes.probe_scene()
    eng = scenes.make_engine(sc, max_particles=6000)
    orc = scenes.make_oracle(sc)
    eng.set_particles(sc["pos"], sc["vel"])
    orc.set_particles(sc["pos"], sc["vel"])
    dt = sc["dt"]

    def run(steps):
        for step in range(steps):
            orc.substep_pcisph(dt)
            eng.step_pcisph(dt)
            cc, co = eng.export_cells()
            assert np.array_equal(cc, orc.arr("cell_count")) and np.array_equal(co, orc.arr("cell_order")), f"step {step}"
            cnt, ids = eng.export_neighbors()
            assert np.array_equal(cnt, orc.a["nbr_count"]) and np.array_equal(ids, orc.a["nbr_ids"]), f"step {step}"
            pos, vel = scenes.f32(orc.a["pos"]), scenes.f32(orc.a["vel"])
            orc.a["pos"][:] = pos
            orc.a["vel"][:] = vel
            eng.overwrite_state(pos, vel)

    run(12)
    add_pos, add_vel = scenes.f32(g["add_pos"]), scenes.f32(g["add_vel"])
    for shift in (np.zeros(3), np.array([0.3, 0.0, 0.2])):
        p = scenes.f32(add_pos + shift)
        n0 = eng.n
        eng.append_particles(p, add_vel)
        orc.append_particles(p, add_vel)
        assert eng.n == n0 + len(p) == orc.S.n
        cc, co = eng.export_cells()                     # chains right after the append
        assert np.array_equal(cc, orc.arr("cell_count")) and np.array_equal(co, orc.arr("cell_order"))
        assert np.array_equal(eng.download(bb.POSITION, np.float32)[n0:], p)   # ids are append order
        tr = orc.trace_pcisph(dt)                       # the sub-step right after it
        eng.step_pcisph(dt)
        cc, co = eng.export_cells()
        assert np.array_equal(cc, tr["cell_count"]) and np.array_equal(co, tr["cell_order"])
        cnt, ids = eng.export_neighbors()
        assert np.array_equal(cnt, tr["nbr_count"]) and np.array_equal(ids, tr["nbr_ids"])
        assert np.abs(eng.download(bb.DENSITY) - tr["density"]).max() / 1000.0 < TOL_RHO
        ext = float(np.max(sc["domain_max"] - sc["domain_min"]))
        assert np.abs(eng.download(bb.POSITION) - tr["pos_out"]).max() / ext < TOL_POS
        pos, vel = scenes.f32(orc.a["pos"]), scenes.f32(orc.a["vel"])
        orc.a["pos"][:] = pos
        orc.a["vel"][:] = vel
        eng.overwrite_state(pos, vel)
        run(10)
    assert eng.stats().nan_count == 0


def test_pseudo_viscosity_smoothing_matches_oracle():
    """The cold branch at the end of every sub-step (pseudoViscosity * dt > 0.1; sph_equations3.cpp:341-382, 469-483;
    k_pseudo_aggregate / k_pseudo_interpolate), oracle side pinned bit-exactly against the reference
    (tests/golden/pseudo_run.npz): 10 sub-steps with the state resynced after each, velocities within tolerance."""
    sc = scenes.probe_scene()
    eng = scenes.make_engine(sc, pseudo_viscosity=200.0)
    orc = scenes.make_oracle(sc)
    orc.P.pseudo_viscosity = 200.0
    eng.set_particles(sc["pos"], sc["vel"])
    orc.set_particles(sc["pos"], sc["vel"])
    dt = sc["dt"]
    ext = float(np.max(sc["domain_max"] - sc["domain_min"]))
    for step in range(10):
        orc.substep_pcisph(dt)
        eng.step_pcisph(dt)
        v, vo = eng.download(bb.VELOCITY), orc.a["vel"]
        assert np.abs(v - vo).max() / np.abs(vo).max() < 10 * TOL_VEL, f"step {step}"
        assert np.abs(eng.download(bb.POSITION) - orc.a["pos"]).max() / ext < TOL_POS, f"step {step}"
        pos, vel = scenes.f32(orc.a["pos"]), scenes.f32(orc.a["vel"])
        orc.a["pos"][:] = pos
        orc.a["vel"][:] = vel
        eng.overwrite_state(pos, vel)
