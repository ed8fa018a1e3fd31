#!/usr/bin/env python
"""Run under torchrun on N GPUs: the NCCL slab path against the single-domain engine (rank 0 runs both).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/slab_nccl_check.py

Checks after S sub-steps of a scene whose particles cross the cuts: positions / velocities / densities
bit-identical (FP32) to the single-domain engine, cell counts + chain order identical, particle count
conserved.  Prints one JSON line on rank 0; exit code 1 on mismatch.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import bubbles_b200 as bb  # noqa: E402
import scenes  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    steps = int(os.environ.get("BBX_CHECK_STEPS", "60"))
    n_target = float(os.environ.get("BBX_CHECK_PARTICLES", "2e5"))
    sc = scenes.dam_break_scene(n_target=n_target, jitter=0.0)
    sc["vel"][:] = scenes.f32([1.0, -6.0, 3.0])  # slosh along z so that particles migrate between slabs
    grid = bb.UtilBuildGridForDomain(sc["domain_min"], sc["domain_max"], sc["spacing"], sc["scale"])
    zb, hist = bench.plan_for_ranks(grid, sc["pos"], world)
    cap, gcap = bb.slab_capacity(hist, zb, rank, slack=2.0)
    slab = bb.NcclSlab(grid, sc["spacing"], sc["scale"], zb, rank, world, bench.broadcast_bytes, cap, gcap, device=local)
    eng = slab.engine
    eng.set_colliders(scenes.engine_colliders(sc))
    eng.set_particles_ids(sc["pos"], sc["vel"])
    n0 = eng.n
    # two thirds of the way, then the z cuts are re-planned from the per-plane histogram (bbx_plane_counts summed over the
    # ranks -> bbx_slab_plan -> bbx_slab_plan_step -> bbx_rebalance: planes change hands over NCCL), then the rest
    first = (2 * steps) // 3
    eng.step_many(sc["dt"], first)
    hist_now = torch.from_numpy(eng.plane_counts()).cuda()
    dist.all_reduce(hist_now, op=dist.ReduceOp.SUM)
    target = bb.plan_slabs(hist_now.cpu().numpy(), world)
    zb_now, moves = list(zb), 0
    for _ in range(8):
        if zb_now == list(target):
            break
        step, _ = bb.plan_step(zb_now, target)
        if step == zb_now:
            break
        eng.rebalance(step)
        zb_now = step
        moves += 1
    n_mid = eng.n
    eng.step_many(sc["dt"], steps - first)
    parts = {}
    ids, parts["pos"] = eng.download_owned(bb.POSITION, np.float32)
    _, parts["vel"] = eng.download_owned(bb.VELOCITY, np.float32)
    _, parts["rho"] = eng.download_owned(bb.DENSITY, np.float32)
    cc, co = eng.export_cells()
    st = eng.stats()
    gathered = [None] * world
    dist.gather_object((ids, parts, cc, co, n0, eng.n, st.nan_count, st.ghosts, n_mid), gathered if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        one = scenes.make_engine(sc, device=local)
        one.set_particles(sc["pos"], sc["vel"])
        one.step_many(sc["dt"], steps)
        n = len(sc["pos"])
        res = {"n_gpus": world, "particles": n, "steps": steps, "z_bounds": zb, "z_bounds_rebalanced": zb_now, "rebalance_calls": moves,
               "owned_after_rebalance": [g[8] for g in gathered]}
        owned = np.zeros(n, dtype=np.int32)
        merged = {k: np.zeros((n, 3) if k != "rho" else n, dtype=np.float32) for k in parts}
        ccs = np.zeros(grid.total, dtype=np.int32)
        order = []
        for ids_r, parts_r, cc_r, co_r, *_ in gathered:
            owned[ids_r] += 1
            for k in merged:
                merged[k][ids_r] = parts_r[k]
            ccs += cc_r
            order.append(co_r)
        res["every_particle_owned_once"] = bool((owned == 1).all())
        res["pos_bit_identical"] = bool(np.array_equal(merged["pos"], one.download(bb.POSITION, np.float32)))
        res["vel_bit_identical"] = bool(np.array_equal(merged["vel"], one.download(bb.VELOCITY, np.float32)))
        res["density_bit_identical"] = bool(np.array_equal(merged["rho"], one.download(bb.DENSITY, np.float32)))
        c1, o1 = one.export_cells()
        res["cell_counts_identical"] = bool(np.array_equal(ccs, c1))
        res["chain_order_identical"] = bool(np.array_equal(np.concatenate(order), o1))
        res["owned_before"] = [g[4] for g in gathered]
        res["owned_after"] = [g[5] for g in gathered]
        res["migrated"] = res["owned_before"] != res["owned_after"]
        res["ghosts"] = [g[7] for g in gathered]
        res["nan"] = int(sum(g[6] for g in gathered))
        ok = all(res[k] for k in ("every_particle_owned_once", "pos_bit_identical", "vel_bit_identical",
                                  "density_bit_identical", "cell_counts_identical", "chain_order_identical")) and res["nan"] == 0
        res["ok"] = ok
        print(json.dumps(res), flush=True)
    eng.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
