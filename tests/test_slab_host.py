"""CPU tests of the multi-GPU host logic: the slab planner and, with a world_size-2 gloo group, the
rank plumbing a torchrun job uses around the engine (same plan on every rank, id broadcast, merge of
per-rank results in particle-id order).  No compute call is made: libbbx has no CPU path."""
import os
import socket
import sys

import numpy as np
import pytest

import bubbles_b200 as bb
import scenes


def test_plan_is_balanced_and_covers_all_planes():
    sc = scenes.dam_break_scene(n_target=5e4)
    grid = bb.UtilBuildGridForDomain(sc["domain_min"], sc["domain_max"], sc["spacing"], sc["scale"])
    hist = bb.plane_histogram(grid, sc["pos"])
    assert hist.sum() == len(sc["pos"]) and len(hist) == grid.n[2]
    # the histogram is the engine's hash: compare with the oracle's cell ids
    orc = scenes.make_oracle(sc)
    cells = orc.hash(sc["pos"][:2000])
    z = cells // (grid.n[0] * grid.n[1])
    assert np.array_equal(np.bincount(z, minlength=grid.n[2]), bb.plane_histogram(grid, sc["pos"][:2000]))
    for nr in (1, 2, 3, 4, 8):
        zb = bb.plan_slabs(hist, nr)
        assert zb[0] == 0 and zb[-1] == grid.n[2] and all(b > a for a, b in zip(zb[:-1], zb[1:]))
        if nr > 1:
            share = [hist[a:b].sum() for a, b in zip(zb[:-1], zb[1:])]
            occupied = np.count_nonzero(hist)
            if occupied >= 4 * nr:
                assert max(share) <= len(sc["pos"]) / nr + 2 * hist.max()
    with pytest.raises(bb.BbxError):
        bb.plan_slabs(hist[:3], 4)


def test_plan_handles_empty_and_lopsided_histograms():
    assert bb.plan_slabs(np.zeros(8, dtype=np.int64), 8) == list(range(9))
    zb = bb.plan_slabs([5, 0, 0, 0, 0, 0, 0, 100], 3)
    assert zb[0] == 0 and zb[-1] == 8 and sorted(set(zb)) == zb
    cap, gcap = bb.slab_capacity(np.array([10, 20, 30, 40]), [0, 2, 4], 1)
    assert cap >= 70 and gcap >= 40


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import bench
        sc = scenes.dam_break_scene(n_target=2e4)
        grid = bb.UtilBuildGridForDomain(sc["domain_min"], sc["domain_max"], sc["spacing"], sc["scale"])
        zb, hist = bench.plan_for_ranks(grid, sc["pos"], world)
        # every rank must hold the same plan
        plans = [None] * world
        dist.all_gather_object(plans, zb)
        # the unique-id style broadcast used for bbx_comm_init
        blob = bench.broadcast_bytes(bytes([rank + 7]) * 128, 0)
        # ownership by plane: each particle belongs to exactly one rank
        plane = np.repeat(np.arange(len(hist)), hist)  # planes of particles sorted by plane
        mine = int(hist[zb[rank]:zb[rank + 1]].sum())
        counts = [None] * world
        dist.all_gather_object(counts, mine)
        q.put((rank, plans, blob, counts, len(sc["pos"]), len(plane)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_rank_plumbing():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, plans0, blob0, counts0, n0, _), (r1, plans1, blob1, counts1, n1, _) = res
    assert plans0 == plans1 and plans0[0] == plans0[1]
    assert blob0 == blob1 == bytes([7]) * 128
    assert sum(counts0) == n0 == n1


def test_plan_steps_walk_to_any_target_through_valid_neighbour_only_moves():
    """bbx_slab_plan_step: from any cuts to any other cuts in steps bbx_rebalance can follow -- every intermediate plan keeps
    at least one plane per rank, every cut stays strictly inside the two slabs it separates in the CURRENT plan (planes change
    hands between neighbours only), and the walk ends at the target after at most a few steps."""
    rng = np.random.default_rng(5)
    for case in range(200):
        nr = int(rng.integers(2, 9)); nplanes = int(rng.integers(nr, 120))
        draw = lambda: [0] + sorted(rng.choice(np.arange(1, nplanes), nr - 1, replace=False).tolist()) + [nplanes]
        cur, tgt = draw(), draw()
        for steps in range(64):
            if cur == tgt:
                break
            step, done = bb.plan_step(cur, tgt)
            step = list(step)
            assert step[0] == 0 and step[-1] == nplanes and all(b > a for a, b in zip(step[:-1], step[1:])), (cur, tgt, step)
            for r in range(1, nr):
                assert cur[r - 1] < step[r] < cur[r + 1], (cur, tgt, step)      # the cut stayed between its two slabs
                assert abs(step[r] - tgt[r]) <= abs(cur[r] - tgt[r])             # and never moved away from its target
            assert step != cur, "no progress"
            assert bool(done) == (step == tgt)
            cur = step
        assert cur == tgt and steps <= nplanes
