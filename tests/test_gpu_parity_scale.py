"""Config-scale parity gate (BASELINE.json configs[1] and [2]):

* configs[1] -- PCISPH dam break, 1 M particles, one B200: cell order and neighbour lists BIT-EXACT against the oracle
  at sub-steps 0, 1 and 10 from injected (identical, FP32-representable) state, every field of the sub-step within
  the FP32-vs-FP64 tolerances; then on a DEVELOPED flow (300 free sub-steps of the engine, splash and spray) where
  the fallback branches of the kernels provably run (sweep tiles that exceed the shared-memory stage, list groups
  redone with the FP64 predicate).  These are the paths the headline number of bench.py runs through: the
  look-back scan over hundreds of tiles, the persistent list build, the staged sweeps.
* configs[2] -- the 8 M-particle run against a baked-SDF torus (stand-in for the absent whale / dragon meshes,
  SURVEY F10): one traced sub-step on the developed flow, collider response included.

Reference: Grid::DistributeToCellOpt / DistributeParticleBucket src/core/grid.h:422-494, scene
src/tests/test_pcisph_extra.cpp:1102-1169.
"""
import numpy as np
import pytest

import bubbles_b200 as bb
import parity_gate as pg
import scenes

pytestmark = pytest.mark.gpu


def _assert_gate(r, where):
    assert r["cell_counts_bit_exact"] and r["cell_order_bit_exact"], f"{where}: cell order differs {r}"
    assert r["neighbor_counts_bit_exact"] and r["neighbor_lists_bit_exact"], f"{where}: neighbour lists differ {r}"
    assert r["overflow_matches"] and r["rebuild_flag_matches"], f"{where}: {r}"
    assert r["fields_within_tolerance"], f"{where}: {r}"
    assert r["nan_count"] == 0


def test_dam_break_1m_lists_bit_exact_at_substeps_0_1_10_and_on_developed_flow():
    sc = scenes.dam_break_scene(1.0e6, jitter=0.0)      # the bench.py workload (configs[1])
    n = len(sc["pos"])
    assert 0.99e6 < n < 1.02e6
    ext = float(np.max(sc["domain_max"] - sc["domain_min"]))
    dt = sc["dt"]
    eng = scenes.make_engine(sc)
    orc = scenes.make_oracle(sc)
    eng.set_particles(sc["pos"], sc["vel"])
    orc.set_particles(sc["pos"], sc["vel"])
    assert eng.stats().occupied_cells > 148 * 5 * 4   # more cells than the persistent list grid has warps: grid-stride path
    # sub-step 0 (incremental update right after Setup's ascending-id rebuild), then 1, then 10
    r, _ = pg.gate_substep(eng, orc, dt, ext)
    _assert_gate(r, "sub-step 0")
    done = 1
    for k in (1, 10):
        while done < k:
            orc.substep_pcisph(dt)
            done += 1
        pg.sync_engine_from_oracle(eng, orc)
        r, _ = pg.gate_substep(eng, orc, dt, ext)
        done += 1
        _assert_gate(r, f"sub-step {k}")
    # developed flow: the engine runs free (FP32), the oracle restarts from its state
    eng.step_many(dt, 300)
    assert eng.stats().nan_count == 0
    pg.sync_oracle_from_engine(eng, orc)
    r, tr = pg.gate_substep(eng, orc, dt, ext)
    _assert_gate(r, "developed flow (300 free sub-steps)")
    # the branches the pristine lattice never takes
    assert r["exact_passes"] > 0, "no list group went through the FP64 IsWithinStd re-check"
    assert r["unstaged_tiles"] > 0, "no sweep tile exceeded the shared-memory stage (fallback branch untested)"
    assert int(tr["cell_count"].max()) > 16, "flow not developed: no compressed cells"
    eng.close()


def test_sdf_torus_8m_traced_substep_on_developed_flow():
    sc = scenes.dam_break_scene_slab(8.0e6, 0, 1, obstacle=scenes.torus_obstacle)
    n = len(sc["pos"])
    assert 7.9e6 < n < 8.2e6
    ext = float(np.max(sc["domain_max"] - sc["domain_min"]))
    dt = sc["dt"]
    eng = scenes.make_engine(sc)
    orc = scenes.make_oracle(sc)
    eng.set_particles(sc["pos"], sc["vel"])
    orc.set_particles(sc["pos"].astype(np.float64), sc["vel"].astype(np.float64))
    # the front of the collapsing block reaches the torus within ~100 sub-steps
    eng.step_many(dt, 160)
    assert eng.stats().nan_count == 0
    pg.sync_oracle_from_engine(eng, orc)
    pos_in = orc.a["pos"].copy()
    r, tr = pg.gate_substep(eng, orc, dt, ext)
    _assert_gate(r, "8 M + SDF torus, sub-step 160")
    # the SDF collider really took part: particles whose integration was redirected next to the torus
    torus = sc["colliders"][1]
    moved = np.abs(tr["pos_out"] - (pos_in + dt * tr["vel_out"])).max(axis=1) > 1e-9
    near = torus["sdf"](tr["pos_out"]) < 2.5 * sc["spacing"]
    assert int((moved & near).sum()) > 0, "no particle collided with the SDF torus"
    eng.close()
