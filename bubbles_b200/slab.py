"""Host side of the multi-GPU z-slab decomposition (include/bbx.h, "multi-GPU").

The reference is single-GPU; what a multi-GPU host has to do around the engine is:
plan the slabs (whole cell planes, balanced by particle count), create one slab engine per rank, attach
the communicator, hand the particles over, and merge per-rank results back into particle-id order.

  plan_slabs / plane_histogram   the planner (bbx_slab_plan / bbx_plane_histogram)
  NcclSlab                        one rank of a torchrun / MPI-style job: one process per GPU, NCCL
  LocalSlabGroup                  all slabs in ONE process on ONE device, one host thread per slab
                                  (bbx_comm_init_local): the same engine code path with copies instead of
                                  NCCL, used to verify the slab logic bit-exactly on a single GPU
"""
import ctypes as C
import threading
import uuid

import numpy as np

from . import _lib as L
from .engine import Engine, _check


def plane_histogram(grid, pos):
    pos = np.ascontiguousarray(pos)
    if pos.dtype != np.float32:
        pos = np.ascontiguousarray(pos, dtype=np.float64)
    out = (C.c_longlong * grid.n[2])()
    _check(L.load().bbx_plane_histogram(C.byref(grid), len(pos), pos.ctypes.data,
                                        L.F32 if pos.dtype == np.float32 else L.F64, out))
    return np.array(out[:], dtype=np.int64)


def plan_slabs(plane_counts, nranks):
    pc = (C.c_longlong * len(plane_counts))(*[int(x) for x in plane_counts])
    zb = (C.c_int * (nranks + 1))()
    _check(L.load().bbx_slab_plan(len(plane_counts), pc, nranks, zb))
    return list(zb[:])


def plan_step(current, target):
    """(step, done): the part of the way from the cuts `current` to `target` that one bbx_rebalance can go
    (bbx_slab_plan_step: every cut stays strictly inside the two slabs it separates)"""
    nr = len(current) - 1
    cur = (C.c_int * (nr + 1))(*[int(z) for z in current])
    tgt = (C.c_int * (nr + 1))(*[int(z) for z in target])
    step = (C.c_int * (nr + 1))()
    done = C.c_int()
    _check(L.load().bbx_slab_plan_step(nr, cur, tgt, step, C.byref(done)))
    return list(step[:]), bool(done.value)


def slab_capacity(plane_counts, z_bounds, rank, slack=1.5, floor=4096):
    """(max_particles, ghost_capacity) of one slab: its share of the particles with room for migration."""
    z0, z1 = z_bounds[rank], z_bounds[rank + 1]
    own = int(np.sum(plane_counts[z0:z1]))
    mx = int(np.max(plane_counts)) if len(plane_counts) else 0
    return max(floor, int(own * slack) + mx), max(floor, int(mx * 2))


class NcclSlab:
    """One rank of a one-process-per-GPU job.  `broadcast_bytes(b, src)` must return rank `src`'s bytes on
    every rank (e.g. torch.distributed.broadcast_object_list) -- used once, for the NCCL unique id."""

    def __init__(self, grid, spacing, kernel_scale, z_bounds, rank, nranks, broadcast_bytes, max_particles,
                 ghost_capacity=0, device=0, **kw):
        self.rank, self.nranks, self.grid = rank, nranks, grid
        self.engine = Engine(grid, spacing, kernel_scale, max_particles, device=device,
                             slab=(z_bounds[rank], z_bounds[rank + 1]) if nranks > 1 else None,
                             ghost_capacity=ghost_capacity, **kw)
        if nranks > 1:
            uid = (C.c_ubyte * 128)()
            if rank == 0:
                _check(self.engine.lib.bbx_comm_unique_id(uid))
            uid = broadcast_bytes(bytes(uid), 0)
            self.engine.comm_init(rank, nranks, uid)


class LocalSlabGroup:
    """nslabs slab engines on one device, stepped in lock-step by one host thread each."""

    def __init__(self, grid, spacing, kernel_scale, z_bounds, max_particles, ghost_capacity=0, device=0, **kw):
        self.grid = grid
        self.z_bounds = list(z_bounds)
        self.nslabs = len(z_bounds) - 1
        name = "bbx-local-" + uuid.uuid4().hex
        caps = max_particles if isinstance(max_particles, (list, tuple)) else [max_particles] * self.nslabs
        self.engines = [Engine(grid, spacing, kernel_scale, caps[r], device=device,
                               slab=(z_bounds[r], z_bounds[r + 1]) if self.nslabs > 1 else None,
                               ghost_capacity=ghost_capacity, **kw) for r in range(self.nslabs)]
        if self.nslabs > 1:
            for r, e in enumerate(self.engines):
                e.comm_init_local(r, self.nslabs, name)

    def close(self):
        for e in self.engines:
            e.close()

    def each(self, fn):
        """fn(engine, rank) on one thread per slab (the engine calls are collective); returns the results."""
        out, err = [None] * self.nslabs, [None] * self.nslabs

        def run(r):
            try:
                out[r] = fn(self.engines[r], r)
            except BaseException as ex:  # noqa: BLE001 -- re-raised below
                err[r] = ex
        th = [threading.Thread(target=run, args=(r,)) for r in range(self.nslabs)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for ex in err:
            if ex is not None:
                raise ex
        return out

    # -- the Engine surface, group-wide
    def set_colliders(self, colliders):
        for e in self.engines:
            e.set_colliders(colliders)

    def set_particles(self, pos, vel, ids=None):
        self.n_total = len(pos)
        self.each(lambda e, r: e.set_particles_ids(pos, vel, ids))

    def append_particles(self, pos, vel):
        """ContinuousParticleSetBuilder3::AddParticle + Commit on the group: ids continue from the global count, every
        slab keeps the particles of its planes"""
        first = sum(self.counts)
        ids = np.arange(first, first + len(pos), dtype=np.int32)
        self.each(lambda e, r: e.append_particles_ids(pos, vel, ids))
        self.n_total = first + len(pos)

    def rebalance(self, z_bounds=None, max_calls=8):
        """Move the cuts to z_bounds (default: bbx_slab_plan over the current plane histogram), in as many neighbour-only
        steps as it takes; returns the cuts reached."""
        if self.nslabs < 2:
            return self.z_bounds
        if z_bounds is None:
            hist = sum(e.plane_counts() for e in self.engines)
            z_bounds = plan_slabs(hist, self.nslabs)
        for _ in range(max_calls):
            if list(z_bounds) == list(self.z_bounds):
                break
            step, done = plan_step(self.z_bounds, z_bounds)
            if step == list(self.z_bounds):
                break
            self.each(lambda e, r: e.rebalance(step))
            self.z_bounds = step
        return self.z_bounds

    def step_pcisph(self, dt):
        self.each(lambda e, r: e.step_pcisph(dt))

    def step_sph(self, dt):
        self.each(lambda e, r: e.step_sph(dt))

    def step_many(self, dt, n, solver=L.SOLVER_PCISPH):
        self.each(lambda e, r: e.step_many(dt, n, solver))

    def advance(self, seconds, solver=L.SOLVER_PCISPH):
        return self.each(lambda e, r: e.advance(seconds, solver))

    def run_phase(self, phase, dt):
        self.each(lambda e, r: e.run_phase(phase, dt))

    def stats(self):
        return [e.stats() for e in self.engines]

    @property
    def counts(self):
        return [e.n for e in self.engines]

    def download(self, field, dtype=np.float64):
        """Merged over the slabs, in particle-id order."""
        parts = [e.download_owned(field, dtype) for e in self.engines]
        n = sum(len(i) for i, _ in parts)
        first = next(v for _, v in parts if v is not None)
        out = np.zeros((n,) + first.shape[1:], dtype=first.dtype)
        seen = np.zeros(n, dtype=bool)
        for ids, v in parts:
            assert not seen[ids].any(), "a particle is owned by two slabs"
            seen[ids] = True
            out[ids] = v
        assert seen.all(), "a particle is owned by no slab"
        return out

    def export_cells(self):
        """(cell_count over the global grid, cell_order): slabs own disjoint, ascending ranges of cell ids, so
        the global order is the concatenation of theirs."""
        cc = np.zeros(self.grid.total, dtype=np.int32)
        order = []
        for e in self.engines:
            c, o = e.export_cells()
            cc += c
            order.append(o)
        return cc, np.concatenate(order)

    def export_neighbors(self):
        n = sum(self.counts)
        counts = np.zeros(n, dtype=np.int32)
        nbr = np.full((n, L.MAX_NEIGHBORS), -1, dtype=np.int32)
        for e in self.engines:
            ids, _ = e.download_owned(None)
            c, i = e.export_neighbors_owned()
            counts[ids] = c
            nbr[ids] = i
        return counts, nbr
