"""GPU debugging aid: the slab append test step by step, first difference in detail."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes
import bubbles_b200 as bb
import test_gpu_slabs as T
from oracle import oracle as O
sc = T._moving_scene()
n0 = len(sc["pos"])
pts = scenes.f32(O.bcc_points((-0.12, 0.12, -0.14), (0.1, 0.2, 0.16), 0.02))
vel = scenes.f32(np.tile([0.3, -2.0, 0.5], (len(pts), 1)))
cap = n0 + 2 * len(pts)
one = T._single(sc, max_particles=cap)
grp, zb = T._group(sc, 3, cap=cap)
grp.set_particles(sc["pos"], sc["vel"])
dt = sc["dt"]
def cmp(tag):
    d1, dg = one.download(bb.DENSITY, np.float32), grp.download(bb.DENSITY, np.float32)
    p1, pg_ = one.download(bb.POSITION, np.float32), grp.download(bb.POSITION, np.float32)
    bad = np.nonzero(d1 != dg)[0]
    badp = np.nonzero((p1 != pg_).any(axis=1))[0]
    print(tag, "density diffs", len(bad), "position diffs", len(badp), flush=True)
    if len(bad):
        n1, i1 = one.export_neighbors(); ng, ig = grp.export_neighbors()
        print("  counts equal", np.array_equal(n1, ng), "lists equal", np.array_equal(i1, ig))
        for i in bad[:6]:
            print("  particle", i, "rho", d1[i], dg[i], "cnt", n1[i], ng[i], "lists same", np.array_equal(i1[i], ig[i]))
        print("  stats one", one.stats().neighbor_overflow, one.stats().max_candidates, "grp", [(s.neighbor_overflow, s.max_candidates) for s in grp.stats()])
        return True
    return False
for k in range(15):
    one.step_pcisph(dt); grp.step_pcisph(dt)
cmp("pre")
done = False
for rep, shift in enumerate((np.zeros(3), np.array([0.01, 0.0, -0.01]))):
    p = scenes.f32(pts + shift)
    one.append_particles(p, vel); grp.append_particles(p, vel)
    for step in range(12):
        one.step_pcisph(dt); grp.step_pcisph(dt)
        if cmp(f"rep {rep} step {step}"):
            done = True; break
    if done: break
