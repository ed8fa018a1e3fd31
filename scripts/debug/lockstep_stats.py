"""CPU estimate (oracle as data source): how many warp steps a run-lockstep list walk needs per tile of 32 slots,
against the per-lane walk (max list length of the tile)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import scenes
n = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0e5
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 30
sc = scenes.dam_break_scene(n)
orc = scenes.make_oracle(sc)
orc.set_particles(sc["pos"], sc["vel"])
for _ in range(warm):
    orc.substep_pcisph(sc["dt"])
tr = orc.trace_pcisph(sc["dt"])
g = orc.grid
cnt, ids = tr["nbr_count"], tr["nbr_ids"]
order = tr["cell_order"]            # slot -> particle id
N = len(order)
pos = tr["pos_in"] if "pos_in" in tr else None
cell_of = np.zeros(N, dtype=np.int64)
cc = tr["cell_count"]
cell_of[order] = np.repeat(np.arange(len(cc)), cc)
nx, ny = int(g["n"][0]), int(g["n"][1])
def cyz(c):
    return (c // nx) % ny, c // (nx * ny)
rows = ids.shape[1]
valid = np.arange(rows)[None, :] < cnt[:, None]
nb = np.where(valid, ids, 0)
cy_i, cz_i = cyz(cell_of); cy_j, cz_j = cyz(cell_of[nb])
run = (cy_j - cy_i[:, None] + 1) * 3 + (cz_j - cz_i[:, None] + 1)
per_run = np.stack([(valid & (run == r)).sum(1) for r in range(9)], 1)   # particle x 9
per_run_slot = per_run[order]
cnt_slot = cnt[order]
T = (N + 31) // 32
pad = T * 32 - N
pr = np.concatenate([per_run_slot, np.zeros((pad, 9), int)]).reshape(T, 32, 9)
cs = np.concatenate([cnt_slot, np.zeros(pad, int)]).reshape(T, 32)
lock = pr.max(1).sum(1)        # steps of the lockstep walk per tile
lane = cs.max(1)               # steps of the per-lane walk
print("particles", N, "mean list", cnt.mean(), "per-run mean", per_run.mean(0).round(2))
print("tile steps: lockstep mean %.1f  per-lane (max cnt) mean %.1f  useful mean %.1f" % (lock.mean(), lane.mean(), cs.mean()))
print("per-run max over tile, mean:", pr.max(1).mean(0).round(2))
