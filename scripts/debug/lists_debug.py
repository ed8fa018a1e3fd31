"""GPU debugging aid: engine vs oracle neighbour lists on small scenes, with the first mismatches printed."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes, parity_gate as pg
import bubbles_b200 as bb

def run(sc, steps, label):
    eng = scenes.make_engine(sc); orc = scenes.make_oracle(sc)
    eng.set_particles(sc["pos"], sc["vel"]); orc.set_particles(sc["pos"], sc["vel"])
    ext = float(np.max(sc["domain_max"] - sc["domain_min"]))
    for k in range(steps):
        pg.sync_engine_from_oracle(eng, orc)
        tr = orc.trace_pcisph(sc["dt"])
        eng.run_phase(bb.PHASE_GRID, sc["dt"])
        eng.run_phase(bb.PHASE_DENSITY, sc["dt"])
        cnt, ids = eng.export_neighbors()
        rho = eng.download(bb.DENSITY)
        badc = np.nonzero(cnt != tr["nbr_count"])[0]
        badl = np.nonzero((ids != tr["nbr_ids"]).any(axis=1))[0]
        erho = np.abs(rho - tr["density"]).max() / 1000.0
        print(label, "step", k, "n", len(cnt), "count mismatches", len(badc), "list mismatches", len(badl), "err_rho", erho, "stats", eng.stats().exact_passes, eng.stats().neighbor_overflow, tr["overflow"], flush=True)
        for i in badl[:5]:
            print("  particle", i, "engine cnt", cnt[i], "oracle cnt", tr["nbr_count"][i])
            print("   engine", ids[i][:cnt[i]].tolist())
            print("   oracle", tr["nbr_ids"][i][:tr["nbr_count"][i]].tolist())
        for ph in (bb.PHASE_FORCE_NP, bb.PHASE_PRESSURE, bb.PHASE_PRESSURE_FORCE, bb.PHASE_INTEGRATE):
            eng.run_phase(ph, sc["dt"])
        orc.a["pos"][:] = tr["pos_out"]; orc.a["vel"][:] = tr["vel_out"]
    eng.close()

if __name__ == "__main__":
    run(scenes.probe_scene(), 3, "probe")
    sc = scenes.block_scene((0.6, 0.6, 0.6), (0.5, 0.12, 0.5), (0.0, -0.2, 0.0), (0, -1, 0))
    run(sc, 2, "slab-like")
