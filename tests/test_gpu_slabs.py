"""Multi-slab parity on ONE GPU: the z-slab decomposition (ghost planes, migration, per-phase halo
exchanges, global flags) driven through the in-process transport (bbx_comm_init_local) must reproduce the
single-domain engine BIT FOR BIT -- same kernels, same chain order, same summation order -- and therefore
the oracle's cell orderings and neighbour lists.  The NCCL transport differs only in how the same byte
ranges travel (bench.py --gpus N exercises it)."""
import numpy as np
import pytest

import bubbles_b200 as bb
import scenes

pytestmark = pytest.mark.gpu


def _group(sc, nslabs, cap=None, **kw):
    grid = bb.UtilBuildGridForDomain(sc["domain_min"], sc["domain_max"], sc["spacing"], sc["scale"])
    hist = bb.plane_histogram(grid, sc["pos"])
    zb = bb.plan_slabs(hist, nslabs)
    n = len(sc["pos"])
    g = bb.LocalSlabGroup(grid, sc["spacing"], sc["scale"], zb, cap or n, ghost_capacity=n, **kw)
    g.set_colliders(scenes.engine_colliders(sc))
    return g, zb


def _single(sc, **kw):
    eng = scenes.make_engine(sc, **kw)
    eng.set_particles(sc["pos"], sc["vel"])
    return eng


def _moving_scene():
    # the block falls and sloshes in +z so that particles cross slab boundaries (migration both ways later)
    sc = scenes.block_scene((0.6, 0.6, 0.6), (0.2, 0.3, 0.2), (0.0, -0.1, 0.0), (0.5, -2.0, 1.5))
    return sc


@pytest.mark.parametrize("nslabs", [2, 3, 4])
def test_slabs_match_single_domain_bit_exact(nslabs):
    sc = _moving_scene()
    one = _single(sc)
    grp, zb = _group(sc, nslabs)
    grp.set_particles(sc["pos"], sc["vel"])
    assert sum(grp.counts) == len(sc["pos"])
    cc, co = grp.export_cells()
    c1, o1 = one.export_cells()
    assert np.array_equal(cc, c1) and np.array_equal(co, o1)
    dt = sc["dt"]
    start_counts = grp.counts
    for step in range(60):
        one.step_pcisph(dt)
        grp.step_pcisph(dt)
        if step % 10 == 9 or step < 3:
            cc, co = grp.export_cells()
            c1, o1 = one.export_cells()
            assert np.array_equal(cc, c1), f"cell counts differ at step {step}"
            assert np.array_equal(co, o1), f"chain order differs at step {step}"
            for f in (bb.POSITION, bb.VELOCITY, bb.DENSITY, bb.PRESSURE, bb.FORCE, bb.PRESSURE_FORCE):
                a, b = grp.download(f, np.float32), one.download(f, np.float32)
                assert np.array_equal(a, b), f"field {f} differs at step {step}: {np.abs(a - b).max()}"
    n1, i1 = one.export_neighbors()
    ng, ig = grp.export_neighbors()
    assert np.array_equal(n1, ng) and np.array_equal(i1, ig)
    assert sum(grp.counts) == len(sc["pos"])
    assert grp.counts != start_counts, "no particle migrated between slabs: the scene does not test migration"
    assert all(s.nan_count == 0 for s in grp.stats())
    grp.close()


def test_slabs_match_oracle_lists_and_order():
    sc = _moving_scene()
    grp, zb = _group(sc, 3)
    orc = scenes.make_oracle(sc)
    grp.set_particles(sc["pos"], sc["vel"])
    orc.set_particles(sc["pos"], sc["vel"])
    dt = sc["dt"]
    # first sub-step from identical inputs: bit-exact integer results, FP32-vs-FP64 tolerance on fields
    tr = orc.trace_pcisph(dt)
    grp.step_pcisph(dt)
    cc, co = grp.export_cells()
    assert np.array_equal(cc, tr["cell_count"]) and np.array_equal(co, tr["cell_order"])
    cnt, ids = grp.export_neighbors()
    assert np.array_equal(cnt, tr["nbr_count"]) and np.array_equal(ids, tr["nbr_ids"])
    assert np.abs(grp.download(bb.DENSITY) - tr["density"]).max() / 1000.0 < 2e-5
    ext = float(np.max(sc["domain_max"] - sc["domain_min"]))
    assert np.abs(grp.download(bb.POSITION) - tr["pos_out"]).max() / ext < 1e-6
    # free-running trajectory against the FP64 oracle
    for _ in range(59):
        orc.substep_pcisph(dt)
        grp.step_pcisph(dt)
    d = np.linalg.norm(grp.download(bb.POSITION) - orc.a["pos"], axis=1)
    assert np.quantile(d, 0.999) <= 1e-2 * sc["spacing"]
    grp.close()


def test_slabs_big_move_rule_is_global():
    """One fast particle in one slab must force the full (ascending-id) rebuild in EVERY slab, like the
    reference's single global flag (sph_equations3.cpp:330-335, 517-523)."""
    sc = _moving_scene()
    z = sc["pos"][:, 2]
    fast = int(np.argmax(z))
    sc["vel"][fast] = scenes.f32([0.0, 0.0, 90.0])  # flies off the block: ~0.035 in the first step (pressure brakes it) > 0.9 * cell length 0.036
    one = _single(sc)
    grp, zb = _group(sc, 3)
    grp.set_particles(sc["pos"], sc["vel"])
    dt = sc["dt"]
    for step in range(4):
        one.step_pcisph(dt)
        grp.step_pcisph(dt)
        if step == 0:
            assert one.stats().rebuild_flag == 1
        if step == 1:
            assert one.stats().full_rebuild == 1
            assert all(s.full_rebuild == 1 for s in grp.stats())
        cc, co = grp.export_cells()
        c1, o1 = one.export_cells()
        assert np.array_equal(cc, c1) and np.array_equal(co, o1), f"step {step}"
    assert np.array_equal(grp.download(bb.POSITION, np.float32), one.download(bb.POSITION, np.float32))
    grp.close()


def test_slabs_sph_and_correct_mode_and_advance():
    sc = _moving_scene()
    # SPH (Jacobi) step
    one = _single(sc)
    grp, zb = _group(sc, 2)
    grp.set_particles(sc["pos"], sc["vel"])
    for _ in range(10):
        one.step_sph(1.44e-4)
        grp.step_sph(1.44e-4)
    assert np.array_equal(grp.download(bb.POSITION, np.float32), one.download(bb.POSITION, np.float32))
    assert np.array_equal(grp.download(bb.VELOCITY, np.float32), one.download(bb.VELOCITY, np.float32))
    grp.close()
    # true predict-correct loop: the iteration count is a global decision (max density error over all slabs)
    one = _single(sc, reference_compat=False)
    grp, zb = _group(sc, 2, reference_compat=False)
    grp.set_particles(sc["pos"], sc["vel"])
    for _ in range(12):
        one.step_pcisph(sc["dt"])
        grp.step_pcisph(sc["dt"])
        its = {s.pcisph_iterations for s in grp.stats()}
        assert its == {one.stats().pcisph_iterations}
    assert np.array_equal(grp.download(bb.POSITION, np.float32), one.download(bb.POSITION, np.float32))
    grp.close()
    # CFL sub-stepping: every slab must pick the same dt sequence (global max |f|)
    one = _single(sc)
    grp, zb = _group(sc, 3)
    grp.set_particles(sc["pos"], sc["vel"])
    sub1, _ = one.advance(1.0 / 240.0)
    subs = grp.advance(1.0 / 240.0)
    assert {s for s, _ in subs} == {sub1}
    assert np.array_equal(grp.download(bb.POSITION, np.float32), one.download(bb.POSITION, np.float32))
    grp.close()


def test_slab_with_empty_rank_and_collider_in_ghost_zone():
    """A slab that owns no particle at all still takes part in every exchange; an obstacle sits across a cut."""
    extra = [dict(kind="sphere", radius=0.06, translate=(0.0, -0.27, 0.0), friction=0.2)]
    sc = scenes.block_scene((0.6, 0.6, 0.6), (0.2, 0.3, 0.2), (0.0, -0.1, 0.0), (0.0, -2.0, 0.0), extra_colliders=extra)
    grid = bb.UtilBuildGridForDomain(sc["domain_min"], sc["domain_max"], sc["spacing"], sc["scale"])
    nz = grid.n[2]
    zb = [0, 2, nz // 2, nz]  # planes 0..1 start (and mostly stay) empty; the middle cut goes through the block
    n = len(sc["pos"])
    grp = bb.LocalSlabGroup(grid, sc["spacing"], sc["scale"], zb, n, ghost_capacity=n)
    grp.set_colliders(scenes.engine_colliders(sc))
    grp.set_particles(sc["pos"], sc["vel"])
    assert grp.counts[0] == 0
    one = _single(sc)
    for _ in range(80):
        one.step_pcisph(sc["dt"])
        grp.step_pcisph(sc["dt"])
    assert np.array_equal(grp.download(bb.POSITION, np.float32), one.download(bb.POSITION, np.float32))
    assert np.array_equal(grp.download(bb.VELOCITY, np.float32), one.download(bb.VELOCITY, np.float32))
    grp.close()


def test_slab_overwrite_owned_round_trip():
    """Host run loop on slabs (bench.py's e2e leg): download the owned particles, hand the same rows back with
    bbx_overwrite_owned, step -- must stay bit-identical to the uninterrupted single-domain run; and a change
    made on the host must reach the neighbour's ghost plane."""
    sc = _moving_scene()
    one = _single(sc)
    grp, zb = _group(sc, 3)
    grp.set_particles(sc["pos"], sc["vel"])
    dt = sc["dt"]
    for step in range(12):
        one.step_pcisph(dt)

        def io(e, r):
            ids, p = e.download_owned(bb.POSITION, np.float32)
            _, v = e.download_owned(bb.VELOCITY, np.float32)
            if len(ids) == 0:
                p = np.zeros((0, 3), np.float32); v = np.zeros((0, 3), np.float32)
            e.overwrite_owned(p, v)
            e.step_pcisph(dt)
        grp.each(io)
    assert np.array_equal(grp.download(bb.POSITION, np.float32), one.download(bb.POSITION, np.float32))
    assert np.array_equal(grp.download(bb.VELOCITY, np.float32), one.download(bb.VELOCITY, np.float32))
    # a host-side edit (every velocity zeroed) must act on both sides of the cuts: same as the single domain
    p1 = one.download(bb.POSITION, np.float32)
    one.overwrite_state(p1, np.zeros_like(p1))

    def zero(e, r):
        ids, p = e.download_owned(bb.POSITION, np.float32)
        if len(ids) == 0:
            p = np.zeros((0, 3), np.float32)
        e.overwrite_owned(p, np.zeros_like(p))
    grp.each(zero)
    for _ in range(5):
        one.step_pcisph(dt)
        grp.step_pcisph(dt)
    assert np.array_equal(grp.download(bb.POSITION, np.float32), one.download(bb.POSITION, np.float32))
    assert np.array_equal(grp.download(bb.VELOCITY, np.float32), one.download(bb.VELOCITY, np.float32))
    grp.close()


@pytest.mark.parametrize("p2p", ["1", "0"])
def test_slab_halo_transports_agree(p2p, monkeypatch):
    """The two halo transports -- stores into the neighbours' ghost slots from inside the sweeps (default) and a
    send / recv pair per phase (BBX_P2P=0) -- must both reproduce the single domain bit for bit."""
    monkeypatch.setenv("BBX_P2P", p2p)
    sc = _moving_scene()
    one = _single(sc)
    grp, zb = _group(sc, 3)
    grp.set_particles(sc["pos"], sc["vel"])
    assert all(e.p2p == (p2p == "1") for e in grp.engines)  # (the neighbours' arrays are mapped at the first collective call)
    for _ in range(40):
        one.step_pcisph(sc["dt"])
        grp.step_pcisph(sc["dt"])
    for f in (bb.POSITION, bb.VELOCITY, bb.DENSITY, bb.PRESSURE):
        assert np.array_equal(grp.download(f, np.float32), one.download(f, np.float32)), f"field {f}"
    cc, co = grp.export_cells()
    c1, o1 = one.export_cells()
    assert np.array_equal(cc, c1) and np.array_equal(co, o1)
    grp.close()


def test_append_particles_on_slab_engines_matches_single_domain():
    """SURVEY.md 8(f)2 on slabs: bbx_append_particles_ids = ContinuousParticleSetBuilder3::AddParticle + Commit
    (src/core/grid.h:1409-1441) with the appended block straddling the cuts -- every slab keeps the particles of its
    planes at the tail of their cells' chains (id order), the boundary planes are exchanged again.  Chains, ids and the
    trajectory that follows: bit-identical to the single-domain engine (whose append is pinned against the reference,
    test_append_particles_between_steps_matches_reference_chains)."""
    sc = _moving_scene()
    n0 = len(sc["pos"])
    from oracle import oracle as O
    pts = scenes.f32(O.bcc_points((-0.12, 0.12, -0.14), (0.1, 0.2, 0.16), 0.02))       # a sheet above the block, over all three slabs
    vel = scenes.f32(np.tile([0.3, -2.0, 0.5], (len(pts), 1)))
    cap = n0 + 2 * len(pts)
    one = _single(sc, max_particles=cap)
    grp, zb = _group(sc, 3, cap=cap)
    grp.set_particles(sc["pos"], sc["vel"])
    dt = sc["dt"]

    def same(step):
        cc, co = grp.export_cells()
        c1, o1 = one.export_cells()
        assert np.array_equal(cc, c1) and np.array_equal(co, o1), f"chains differ at {step}"
        for f in (bb.POSITION, bb.VELOCITY, bb.DENSITY):
            assert np.array_equal(grp.download(f, np.float32), one.download(f, np.float32)), f"field {f} differs at {step}"

    for _ in range(15):
        one.step_pcisph(dt)
        grp.step_pcisph(dt)
    for rep, shift in enumerate((np.zeros(3), np.array([0.01, 0.0, -0.01]))):
        p = scenes.f32(pts + shift)
        before = grp.counts
        one.append_particles(p, vel)
        grp.append_particles(p, vel)
        assert sum(grp.counts) == one.n == n0 + (rep + 1) * len(pts)
        assert sum(a != b for a, b in zip(before, grp.counts)) >= 2, "the appended block landed in one slab only"
        same(f"append {rep}")
        for step in range(12):
            one.step_pcisph(dt)
            grp.step_pcisph(dt)
        same(f"after append {rep}")
    n1, i1 = one.export_neighbors()
    ng, ig = grp.export_neighbors()
    assert np.array_equal(n1, ng) and np.array_equal(i1, ig)
    assert all(s.nan_count == 0 for s in grp.stats())
    grp.close()


def test_rebalance_moves_the_cuts_and_stays_bit_identical_to_the_single_domain():
    """bbx_rebalance (include/bbx.h, "re-balancing during a run"): the static plan drifts as the fluid sloshes; the group
    re-plans from the per-plane histogram (bbx_plane_counts -> bbx_slab_plan -> bbx_slab_plan_step) and whole planes change
    hands between neighbours in their chain order.  Also driven far off balance on purpose (cuts pushed several planes
    both ways, more than one neighbour-only step).  Chains, lists and the trajectory: bit-identical to the single domain."""
    sc = _moving_scene()
    n = len(sc["pos"])
    one = _single(sc)
    grp, zb = _group(sc, 3, cap=n)
    grp.set_particles(sc["pos"], sc["vel"])
    dt = sc["dt"]

    def same(tag):
        cc, co = grp.export_cells()
        c1, o1 = one.export_cells()
        assert np.array_equal(cc, c1) and np.array_equal(co, o1), f"chains differ {tag}"
        for f in (bb.POSITION, bb.VELOCITY, bb.DENSITY):
            assert np.array_equal(grp.download(f, np.float32), one.download(f, np.float32)), f"field {f} differs {tag}"

    for _ in range(10):
        one.step_pcisph(dt); grp.step_pcisph(dt)
    same("before")
    counts0 = list(grp.counts)
    # (a) far off balance: the first cut three planes down, the second two planes up -- planes travel both ways
    target = [zb[0], zb[1] - 3, zb[2] + 2, zb[3]]
    reached = grp.rebalance(target)
    assert reached == target and sum(grp.counts) == n and list(grp.counts) != counts0
    hist = sum(e.plane_counts() for e in grp.engines)
    assert hist.sum() == n and [int(hist[reached[r]:reached[r + 1]].sum()) for r in range(3)] == list(grp.counts)
    same("after the forced move")
    for _ in range(12):
        one.step_pcisph(dt); grp.step_pcisph(dt)
    same("12 sub-steps after the forced move")
    # (b) back to balance from the histogram
    reached = grp.rebalance()
    hist = sum(e.plane_counts() for e in grp.engines)
    assert sum(grp.counts) == n and max(grp.counts) - min(grp.counts) <= hist.max()   # (whole planes: ~6 planes hold the block)
    for _ in range(12):
        one.step_pcisph(dt); grp.step_pcisph(dt)
    same("12 sub-steps after re-balancing")
    n1, i1 = one.export_neighbors()
    ng, ig = grp.export_neighbors()
    assert np.array_equal(n1, ng) and np.array_equal(i1, ig)
    # (c) a plan a rank cannot follow stops every rank before anything moves (no hang, no change)
    with pytest.raises(bb.BbxError):
        grp.each(lambda e, r: e.rebalance([0, 1, 2, grp.grid.n[2]]) if reached[1] > 3 else e.rebalance([0, grp.grid.n[2] - 2, grp.grid.n[2] - 1, grp.grid.n[2]]))
    one.step_pcisph(dt); grp.step_pcisph(dt)
    same("after the refused plan")
    assert all(s.nan_count == 0 for s in grp.stats())
    grp.close()
    one.close()
