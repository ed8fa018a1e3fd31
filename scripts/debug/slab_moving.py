"""GPU debugging aid: the slab-cut moving-collider test step by step."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes
import bubbles_b200 as bb
import test_gpu_colliders_moving as T
Z0 = -0.075
sc = T._scene()
sc["colliders"][1]["translate"] = (0.1, -0.25, Z0)
grid = bb.UtilBuildGridForDomain(sc["domain_min"], sc["domain_max"], sc["spacing"], sc["scale"])
zb = bb.plan_slabs(bb.plane_histogram(grid, sc["pos"]), 3)
n = len(sc["pos"])
grp = bb.LocalSlabGroup(grid, sc["spacing"], sc["scale"], zb, n, ghost_capacity=n)
grp.set_colliders(scenes.engine_colliders(sc))
grp.set_particles(sc["pos"], sc["vel"])
one = scenes.make_engine(sc)
one.set_particles(sc["pos"], sc["vel"])
dt = sc["dt"]
for k in range(90):
    c, lin, ang = T._path(k, Z0)
    ec, _ = T._sphere_at(c, lin, ang)
    one.update_collider(1, ec)
    for e in grp.engines:
        e.update_collider(1, ec)
    one.step_pcisph(dt)
    t0 = time.time()
    try:
        grp.step_pcisph(dt)
    except Exception as ex:
        print("step", k, "FAILED after", time.time() - t0, ex, flush=True)
        break
    same = all(np.array_equal(grp.download(f, np.float32), one.download(f, np.float32)) for f in (bb.POSITION, bb.VELOCITY, bb.DENSITY))
    print("step", k, "ok", round(time.time() - t0, 3), "identical" if same else "DIFFERENT", [s.max_candidates for s in grp.stats()], flush=True)
    if not same:
        d = np.abs(grp.download(bb.DENSITY, np.float32) - one.download(bb.DENSITY, np.float32))
        print("  density diffs", (d > 0).sum(), d.max(), np.nonzero(d > 0)[0][:10])
        break
