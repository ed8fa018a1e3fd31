"""The parity gate shared by the config-scale GPU tests and bench.py's "parity" key ("parity gates
run with every benchmark"; SURVEY.md 8(d)-2: neighbour lists compared bit-exact from injected state).

TEST INFRASTRUCTURE: this module drives the oracle (oracle/bbx_oracle.c, pinned bit-exact to the unmodified
reference) as the CHECKER of the CUDA engine.  Nothing here is on the product path or inside a timed region.

  sync_oracle_from_engine   oracle <- the engine's current FP32 state + chains + rebuild flag (identical inputs)
  sync_engine_from_oracle   engine <- the oracle's state rounded to FP32 (on both sides) + chains + flag
  gate_substep              one traced sub-step on both sides; returns bit-exactness flags and max errors
  row_checksums             order-sensitive 64-bit checksum per neighbour-list row (multi-rank gate: rows travel
                            as 8 bytes instead of 400)
"""
import numpy as np

import bubbles_b200 as bb

RHO0 = 1000.0
# the bars of tests/test_gpu_parity.py (engine FP32 vs oracle FP64 from identical inputs, one sub-step)
TOL = dict(rho=2e-5, force=2e-4, pos=1e-6, vel=2e-4, pressure=2e-3, force_p=2e-3)


MAX_FLIP_FRACTION = 2e-6  # particles allowed to take the other branch of a collider-response threshold (see gate_substep)


def f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def sync_oracle_from_engine(eng, orc):
    """Identical inputs for a traced sub-step taken from the ENGINE's current state (e.g. after hundreds of free
    sub-steps): FP32 positions / velocities are exactly representable in the oracle's FP64 arrays."""
    pos, vel = eng.download(bb.POSITION, np.float32), eng.download(bb.VELOCITY, np.float32)
    cc, co = eng.export_cells()
    flag = eng.stats().rebuild_flag
    orc.a["pos"][:] = pos
    orc.a["vel"][:] = vel
    orc.set_chains(cc, co)
    orc.S.rebuild_flag = int(flag)


def sync_engine_from_oracle(eng, orc):
    pos, vel = f32(orc.a["pos"]), f32(orc.a["vel"])
    orc.a["pos"][:] = pos
    orc.a["vel"][:] = vel
    eng.overwrite_state(pos, vel)
    eng.inject_chains(orc.arr("cell_count"), orc.arr("cell_order"))
    eng.set_rebuild_flag(orc.S.rebuild_flag)


_W = None


def row_checksums(ids):
    """ids [n, 100] int32 (unused = -1) -> uint64 [n]: sum_k (ids[k] + 2) * w_k mod 2^64 with fixed odd weights:
    sensitive to content AND position, cheap to gather across ranks."""
    global _W
    if _W is None:
        rng = np.random.default_rng(0xB0B)
        _W = rng.integers(1, 1 << 62, size=ids.shape[1], dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    out = np.zeros(len(ids), dtype=np.uint64)
    step = 1 << 18
    for s in range(0, len(ids), step):
        blk = (ids[s:s + step].astype(np.int64) + 2).astype(np.uint64)
        out[s:s + step] = (blk * _W[None, :]).sum(axis=1, dtype=np.uint64)
    return out


def _relmax(a, b, scale=None):
    scale = scale if scale is not None else max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max() / scale)


def gate_substep(eng, orc, dt, extent, lists=True):
    """One compat PCISPH sub-step phase by phase on both sides from the (identical) current state.  Returns
    (result dict, oracle trace).  result["ok"] = every integer result bit-exact and every field within TOL."""
    tr = orc.trace_pcisph(dt)
    r = {}
    eng.run_phase(bb.PHASE_GRID, dt)
    cc, co = eng.export_cells()
    r["cell_counts_bit_exact"] = bool(np.array_equal(cc, tr["cell_count"]))
    r["cell_order_bit_exact"] = bool(np.array_equal(co, tr["cell_order"]))
    eng.run_phase(bb.PHASE_DENSITY, dt)
    st = eng.stats()
    r["exact_passes"], r["max_candidates"] = int(st.exact_passes), int(st.max_candidates)
    r["neighbor_overflow"] = int(st.neighbor_overflow)
    r["overflow_matches"] = bool(st.neighbor_overflow == tr["overflow"])
    if lists:
        cnt, ids = eng.export_neighbors()
        r["neighbor_counts_bit_exact"] = bool(np.array_equal(cnt, tr["nbr_count"]))
        r["neighbor_lists_bit_exact"] = bool(np.array_equal(ids, tr["nbr_ids"]))
        r["neighbor_entries"] = int(cnt.sum())
        del ids
    else:
        cnt = eng.download(bb.NEIGHBOR_COUNT)
        r["neighbor_counts_bit_exact"] = bool(np.array_equal(cnt, tr["nbr_count"]))
    r["err_density"] = _relmax(eng.download(bb.DENSITY), tr["density"], RHO0)
    eng.run_phase(bb.PHASE_FORCE_NP, dt)
    r["err_force_np"] = _relmax(eng.download(bb.FORCE_NP), tr["force_np"])
    r["err_pos_pred"] = _relmax(eng.download(bb.PRED_POSITION), tr["pos_pred"], extent)
    eng.run_phase(bb.PHASE_PRESSURE, dt)
    r["unstaged_tiles"] = int(eng.stats().unstaged_tiles)
    r["err_density_pred"] = _relmax(eng.download(bb.PRED_DENSITY), tr["density_pred"], RHO0)
    r["err_pressure"] = _relmax(eng.download(bb.PRESSURE), tr["pressure"], max(float(tr["pressure"].max()), 1e-30))
    eng.run_phase(bb.PHASE_PRESSURE_FORCE, dt)
    fscale = max(float(np.abs(tr["force_p"]).max()), float(np.abs(tr["force_np"]).max()), 1e-30)
    r["err_force_p"] = _relmax(eng.download(bb.PRESSURE_FORCE), tr["force_p"], fscale)
    eng.run_phase(bb.PHASE_INTEGRATE, dt)
    r["err_pos"] = _relmax(eng.download(bb.POSITION), tr["pos_out"], extent)
    # velocities: the collider response is DISCONTINUOUS in the position (penetrating iff |sd| < radius, reflect iff
    # v_rel . n < 0), so a particle whose FP32 position sits within rounding of such a threshold can take the other branch
    # than the FP64 oracle: its velocity then differs by (1 + e) v_n while its position (projected either way or not at all
    # by ~0) agrees.  Among millions of particles a handful of such flips is expected; the bar is on all the others.
    dv = np.abs(eng.download(bb.VELOCITY) - tr["vel_out"]).max(axis=1) / max(float(np.abs(tr["vel_out"]).max()), 1e-30)
    flips = dv >= TOL["vel"]
    r["vel_threshold_flips"] = int(flips.sum())
    r["err_vel_max"] = float(dv.max())
    r["err_vel"] = float(dv[~flips].max()) if (~flips).any() else 0.0
    st = eng.stats()
    r["rebuild_flag_matches"] = bool(st.rebuild_flag == tr["rebuild_flag_out"])
    r["nan_count"] = int(st.nan_count)
    r["lists_bit_exact"] = bool(r["cell_counts_bit_exact"] and r["cell_order_bit_exact"] and r["neighbor_counts_bit_exact"]
                                and r.get("neighbor_lists_bit_exact", True) and r["overflow_matches"])
    r["fields_within_tolerance"] = bool(r["err_density"] < TOL["rho"] and r["err_force_np"] < TOL["force"] and r["err_pos_pred"] < TOL["pos"]
                                        and r["err_density_pred"] < TOL["rho"] and r["err_pressure"] < TOL["pressure"]
                                        and r["err_force_p"] < TOL["force_p"] and r["err_pos"] < TOL["pos"] and r["err_vel"] < TOL["vel"]
                                        and r["vel_threshold_flips"] <= max(1, int(MAX_FLIP_FRACTION * len(dv))))
    r["ok"] = bool(r["lists_bit_exact"] and r["fields_within_tolerance"] and r["rebuild_flag_matches"] and r["nan_count"] == 0)
    return r, tr
