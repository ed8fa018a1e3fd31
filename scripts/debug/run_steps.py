"""GPU debugging aid: step a dam break N sub-steps (for compute-sanitizer runs)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes
n = float(sys.argv[1]); steps = int(sys.argv[2])
sc = scenes.dam_break_scene(n)
eng = scenes.make_engine(sc)
eng.set_particles(sc["pos"], sc["vel"])
for k in range(0, steps, 25):
    eng.step_many(sc["dt"], 25)
    eng.synchronize()
    s = eng.stats()
    print(k + 25, "ok", s.exact_passes, s.max_candidates, s.neighbor_overflow, flush=True)
