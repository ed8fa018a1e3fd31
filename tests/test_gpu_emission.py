"""SURVEY.md 8(f)2 on the GPU: continuous re-emission, ContinuousParticleSetBuilder3::MapGrid + MapGridEmit
(src/core/grid.h:1288-1407), with the per-cell occupancy / nearest test on the DEVICE (bbx_query_cells) instead of the
reference's host walk over the chains, and the Commit as bbx_append_particles.

The oracle side (map_grid / map_grid_emit) is pinned bit-exactly against the unmodified reference through a whole run
(tests/golden/emit_run.npz, test_map_grid_emit_bit_exact).  Here the engine follows the oracle with the state resynced
every sub-step (identical FP32-representable inputs): the SAME template positions must be re-emitted in the SAME order,
and the chains after each Commit must match.
"""
import os

import numpy as np
import pytest

import bubbles_b200 as bb
import scenes

pytestmark = pytest.mark.gpu
MAXB = 100


def _engine_map_grid(eng, pos):
    """MapGrid: the chains' positions per occupied cell, in chain order (ascending cell id)"""
    cc, co = eng.export_cells()
    start = np.concatenate([[0], np.cumsum(cc)])
    cells = np.nonzero(cc)[0]
    return cells, [pos[co[start[c]:start[c + 1]]].copy() for c in cells]


def _engine_map_grid_emit(eng, cells, templates, d):
    """MapGridEmit's rule on top of the device query: first min(100 - size, len) template points of every cell with room
    that no particle of the current chain blocks"""
    cid = np.concatenate([np.full(len(t), c, dtype=np.int32) for c, t in zip(cells, templates)])
    pts = np.concatenate(templates)
    size, blocked = eng.query_cells(cid, pts, d)
    out, at = [], 0
    for t in templates:
        sz = int(size[at])
        if 0 <= sz < MAXB:
            k = min(MAXB - sz, len(t))
            keep = ~blocked[at:at + k].astype(bool)
            out.append(t[:k][keep])
        at += len(t)
    return np.concatenate(out) if out else np.zeros((0, 3))


def test_map_grid_emit_on_the_device_follows_the_reference_rule():
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "emit_run.npz"))
    sc = scenes.probe_scene()
    sc["pos"], sc["vel"] = scenes.f32(g["p_pos"]), scenes.f32(g["p_vel"])
    eng = scenes.make_engine(sc, max_particles=8000)
    orc = scenes.make_oracle(sc)
    eng.set_particles(sc["pos"], sc["vel"])
    orc.set_particles(sc["pos"], sc["vel"])
    dt, d = 7e-4, 0.02
    orc.map_grid()
    cells, templates = _engine_map_grid(eng, eng.download(bb.POSITION))
    assert sorted(orc.mapped) == list(cells) and all(np.array_equal(orc.mapped[int(c)], t) for c, t in zip(cells, templates))

    def run(steps):
        for step in range(steps):
            orc.substep_pcisph(dt)
            eng.step_pcisph(dt)
            pos, vel = scenes.f32(orc.a["pos"]), scenes.f32(orc.a["vel"])
            orc.a["pos"][:] = pos
            orc.a["vel"][:] = vel
            eng.overwrite_state(pos, vel)
            cc, co = eng.export_cells()
            assert np.array_equal(cc, orc.arr("cell_count")) and np.array_equal(co, orc.arr("cell_order")), f"step {step}"

    added = []
    for steps in (40, 30):
        run(steps)
        n0 = eng.n
        new = _engine_map_grid_emit(eng, cells, templates, d)
        k = orc.map_grid_emit((0.0, -1.0, 0.0), d)
        assert k == len(new) > 0, (k, len(new))
        assert np.array_equal(orc.a["pos"][n0:], new), "different template positions (or order) re-emitted"
        eng.append_particles(new, np.tile([0.0, -1.0, 0.0], (len(new), 1)))
        assert eng.n == n0 + k == orc.S.n
        cc, co = eng.export_cells()
        assert np.array_equal(cc, orc.arr("cell_count")) and np.array_equal(co, orc.arr("cell_order"))
        added.append(k)
    run(10)
    # the reference's own run (FP64, never rounded) re-emits 301 and 268: the FP32-resynced run must stay close to it
    assert abs(added[0] - int(g["added"][0])) <= 3 and abs(added[1] - int(g["added"][1])) <= 6, (added, g["added"])
    assert eng.stats().nan_count == 0
    eng.close()


def test_query_cells_edge_cases():
    sc = scenes.probe_scene()
    eng = scenes.make_engine(sc)
    # before any particle set: nothing blocks, every cell is empty
    size, blocked = eng.query_cells([0, 5], [[0, 0, 0], [0.1, 0.1, 0.1]], 0.02)
    assert list(size) == [0, 0] and list(blocked) == [0, 0]
    eng.set_particles(sc["pos"], sc["vel"])
    cc, co = eng.export_cells()
    c = int(np.argmax(cc))
    start = int(np.cumsum(cc)[c] - cc[c])
    p = sc["pos"][co[start]]
    size, blocked = eng.query_cells([c, c, -1, eng.grid.total + 7], [p, p + 0.5, p, p], 0.02)
    assert size[0] == size[1] == cc[c] and blocked[0] == 1 and blocked[1] == 0
    assert size[2] == -1 and size[3] == -1            # not a cell of this engine
    with pytest.raises(ValueError):
        eng.query_cells([1, 2], [[0, 0, 0]], 0.02)
    eng.close()
