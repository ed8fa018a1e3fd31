#!/usr/bin/env python
"""bench.py -- particle-updates/s of the 3D PCISPH sub-step (BASELINE.json metric) on N B200s.

One "step" = one PCISPH sub-step (AdvanceTimeStep, reference src/solvers/pcisph_solver3.cpp:42-65)
of the whole synthetic dam-break block; one particle-update = one particle advanced by one sub-step.
  value : state resident in HBM, K sub-steps enqueued back to back on the engine's stream, timed with
          CUDA events on that stream, max over ranks.
  e2e   : the same sub-step through the C ABI with HOST buffers: every step uploads positions +
          velocities from pinned host memory (bbx_overwrite_state), steps, and downloads positions +
          velocities (bbx_download) -- what UtilRunSimulation3 does per frame (src/core/util.h:583-595).
  --impl reference : the unmodified reference's CPU path (oracle/_ref/bbref) on a bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "particle-updates/s (3D PCISPH sub-step, dam break)"
UNIT = "particle-updates/s"
# algorithmic bytes per particle per phase (FP32, vec3 padded to 16 B)
PHASE_BYTES = {"grid": 96, "density": 20, "force_np+predict": 116, "pressure": 24, "pressure_force+integrate": 136}
PHASE_IDS = {"grid": 0, "density": 1, "force_np+predict": 2, "pressure": 4, "pressure_force+integrate": 5}
BYTES_PER_UPDATE = 392  # K = 1 predict-correct iteration (the reference's effective behaviour)
# --solver sph (BASELINE configs[0], SphSolver3): grid 96 + density/EOS 24 + all forces and integrate 72 = 192 B per update
SPH_PHASE_BYTES = {"grid": 96, "density": 24, "forces": 40, "integrate": 32}
SPH_PHASE_IDS = {"grid": 0, "density": 1, "forces": 2, "integrate": 6}
SPH_BYTES_PER_UPDATE = 192


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: an NVML polling thread (every ~2 ms; the timed region of
    a default run lasts only ~50 ms, too short for `nvidia-smi -lms`), falling back to nvidia-smi when NVML is absent."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [x for x in vis.split(",") if x.strip().isdigit()]
        self.index = int(ids[index]) if index < len(ids) else index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []   # (sm_mhz, reasons bitmask)
        self.stop_flag = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag.is_set():
            try:
                self.samples.append((float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)), int(reasons_fn(self.handle))))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nvml:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml:
            self.stop_flag.set()
            self.t.join(timeout=1.0)
            sm = sorted(s for s, _ in self.samples)
            bits = 0
            for _, r in self.samples:
                bits |= r
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(name for bit, name in self.REASONS.items() if bits & bit), "samples": len(sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def workload_label(solver, workload, n_global, n, cells, world):
    """config["workload"]: one string for the bbx arm and the reference arm (same scene -> same label)."""
    return (("PCISPH 3D dam break + baked-SDF torus collider (BASELINE configs[2] stand-in), " if workload == "sdf" else "") +
            f"{'PCISPH' if solver == 'pcisph' else 'SPH'} 3D dam break, {n_global} particles = {n} per GPU (BASELINE configs[{'1' if solver == 'pcisph' else '0'}] per GPU), spacing 0.02, h = 1.8 s, " +
            (f"{cells} cells, fixed dt 7.2e-4, reference-compat (1 predict-correct iteration)" if solver == "pcisph" else
             f"{cells} cells, SphSolver3 step (BASELINE configs[0]), fixed dt 1.44e-4") +
            (f", {world} z-slabs with ghost-plane exchange and migration over NVLink" if world > 1 else ""))


def make_scene(n_target):
    import scenes
    return scenes.dam_break_scene(n_target=n_target, jitter=0.0)


def plan_for_ranks(grid, pos, world):
    """z-slab plan every rank derives identically from the (deterministic, synthetic) scene."""
    import bubbles_b200 as bb
    hist = bb.plane_histogram(grid, pos)
    return bb.plan_slabs(hist, world), hist


def broadcast_bytes(blob, src):
    """rank src's bytes on every rank (used once: the NCCL unique id of the slab group)."""
    import torch.distributed as dist
    box = [blob]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def run_reference(args, rank):
    """Reference arm: unmodified reference CPU path (oracle/_ref/bbref) -- or, if that binary is absent, the
    C port (oracle/liboracle.so) -- on a bounded sample of the workload, all host threads."""
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    n_sample = args.ref_particles
    sc = make_scene(n_sample)
    n = len(sc["pos"])
    steps, warm = max(1, args.steps), max(0, args.warmup)
    # keep the whole run within a few minutes whatever K the driver asks for
    est = n / 1.0e5 / max(1, min(cores, 16))  # s per sub-step, conservative (measured: 3.2e6 updates/s on 16 cores)
    if est * (steps + warm) > 240:
        steps = max(1, int(240 / est) - warm)
    half = sc["domain_max"]
    if O.ref_available():
        kind = "reference"
        wd = tempfile.mkdtemp(prefix="bbref_")
        O.write_particles(os.path.join(wd, "p.bin"), sc["pos"], sc["vel"])
        I = O.mat_str(np.eye(4))
        job = [f"threads {cores}", f"spacing {sc['spacing']}", f"scale {sc['scale']}",
               f"collider box {I} {float(2 * half[0])!r} {float(2 * half[1])!r} {float(2 * half[2])!r} 1 0", "domain_from_collider 0",
               f"particles {wd}/p.bin", "setup"]
        if warm:
            job.append(f"step {sc['dt']} {warm}")
        job.append(f"step {sc['dt']} {steps}")
        out, _ = O.run_ref(job, wd, timeout=3000)
        m = re.findall(r"steps=(\d+) dt=\S+ seconds=(\S+) particle_updates_per_s=(\S+)", out)
        sec = float(m[-1][1]); value = float(m[-1][2])
    else:
        kind = "port"
        orc = __import__("scenes").make_oracle(sc)
        orc.set_particles(sc["pos"], sc["vel"])
        for _ in range(warm):
            orc.substep_pcisph(sc["dt"])
        t0 = time.perf_counter()
        for _ in range(steps):
            orc.substep_pcisph(sc["dt"])
        sec = time.perf_counter() - t0
        value = n * steps / sec
    sample = f"{n}-particle dam break (the bench workload's scene and constants), {steps} sub-steps after {warm} warm-up, FP64, {cores} host threads"
    cells = int(__import__("scenes").make_oracle(sc).grid["total"])
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": 1e3 * sec / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            # N = 1: the very scene the bbx arm steps (same label); N > 1: the 1-GPU scene as the bounded sample of the N-GPU workload
            "config": {"workload": workload_label("pcisph", "dam", n, n, cells, 1), "particles_per_gpu": n, "particles": n, "cells": cells, "dt": sc["dt"],
                       "parallelism": "single",
                       "arm": "unmodified reference, CPU path (AdvanceTimeStep(PciSphSolver3*, dt, use_cpu = 1)), all host threads; the GPU is only touched by the reference's own setup code"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline(n_gpu_particles):
    """Reference CPU path on a bounded sample (10-30 s of CPU work), rank 0, N = 1 only."""
    import numpy as np
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    sc = make_scene(200_000)
    n = len(sc["pos"])
    try:
        if O.ref_available():
            wd = tempfile.mkdtemp(prefix="bbref_")
            O.write_particles(os.path.join(wd, "p.bin"), sc["pos"], sc["vel"])
            half = sc["domain_max"]
            I = O.mat_str(np.eye(4))
            job = [f"threads {cores}", f"spacing {sc['spacing']}", f"scale {sc['scale']}",
                   f"collider box {I} {float(2 * half[0])!r} {float(2 * half[1])!r} {float(2 * half[2])!r} 1 0", "domain_from_collider 0",
                   f"particles {wd}/p.bin", "setup", f"step {sc['dt']} 1", f"step {sc['dt']} 4"]
            out, _ = O.run_ref(job, wd, timeout=900)
            m = re.findall(r"particle_updates_per_s=(\S+)", out)
            return {"value": float(m[-1]), "unit": UNIT, "cores": cores, "kind": "reference",
                    "sample": f"{n}-particle dam break, 4 sub-steps after 1 warm-up, unmodified reference CPU path (FP64), {cores} threads"}
        orc = __import__("scenes").make_oracle(sc)
        orc.set_particles(sc["pos"], sc["vel"])
        orc.substep_pcisph(sc["dt"])
        t0 = time.perf_counter()
        for _ in range(3):
            orc.substep_pcisph(sc["dt"])
        sec = time.perf_counter() - t0
        return {"value": n * 3 / sec, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{n}-particle dam break, 3 sub-steps after 1 warm-up, C port of the reference (FP64, OpenMP)"}
    except Exception as ex:  # the baseline is a reported figure, never a reason to lose the GPU line
        return {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": f"failed: {ex}"}


def reference_gpu(n_particles):
    """The unmodified reference's GPU path (its native mode: managed memory, FP64, 16-thread blocks as shipped) on the
    bench workload itself (capped at 2 M particles), on this box's GPU -- reported beside the CPU path, never the
    optimisation target."""
    import numpy as np
    from oracle import oracle as O
    if not O.ref_gpu_available():
        return {"value": None, "unit": UNIT, "sample": "oracle/_ref/bbref_gpu not built"}
    sc = make_scene(min(n_particles, 2_000_000))
    n = len(sc["pos"])
    try:
        wd = tempfile.mkdtemp(prefix="bbref_gpu_")
        O.write_particles(os.path.join(wd, "p.bin"), sc["pos"], sc["vel"])
        half = sc["domain_max"]
        I = O.mat_str(np.eye(4))
        job = [f"spacing {sc['spacing']}", f"scale {sc['scale']}",
               f"collider box {I} {float(2 * half[0])!r} {float(2 * half[1])!r} {float(2 * half[2])!r} 1 0", "domain_from_collider 0",
               f"particles {wd}/p.bin", "setup", f"step {sc['dt']} 2", f"step {sc['dt']} 10"]
        out, _ = O.run_ref(job, wd, timeout=600, gpu=True)
        m = re.findall(r"particle_updates_per_s=(\S+)", out)
        return {"value": float(m[-1]), "unit": UNIT, "kind": "reference GPU path (unmodified sources, sm_100 build, block size 16 as shipped, FP64, managed memory)",
                "sample": f"{n}-particle dam break, 10 sub-steps after 2 warm-up, on this box's GPU"}
    except Exception as ex:
        return {"value": None, "unit": UNIT, "sample": f"failed: {str(ex)[-300:]}"}


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPU cores next to its GPU BEFORE the pinned host buffers are allocated (first touch puts
    them on that NUMA node): with 8 ranks on one box the host copies otherwise cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = [x for x in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if x.strip().isdigit()]
        h = pynvml.nvmlDeviceGetHandleByIndex(int(vis[index]) if index < len(vis) else index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return True
    except Exception:
        return False


class Job:
    """One engine (single domain or one z-slab of an NCCL group) on this rank's GPU, holding a synthetic scene."""

    def __init__(self, particles_total, workload, rank, world, local_rank):
        import bubbles_b200 as bb
        import scenes
        self.bb, self.rank, self.world = bb, rank, world
        # weak scaling: ONE dam-break scene of (particles per GPU) x N particles, cut into N z-slabs of whole cell planes
        # balanced by particle count; ghost planes / migration / per-phase halos travel over NVLink.  Every rank
        # generates only its own share of the (deterministic) BCC block.
        sc = scenes.dam_break_scene_slab(particles_total, rank, world, obstacle=scenes.torus_obstacle if workload == "sdf" else None)
        self.sc, self.n_global, self.dt = sc, sc["n_global"], sc["dt"]
        if world > 1:
            cap, gcap = bb.slab_capacity(sc["hist"], sc["z_bounds"], rank, slack=2.0)
            self.slab = bb.NcclSlab(sc["grid"], sc["spacing"], sc["scale"], sc["z_bounds"], rank, world, broadcast_bytes, cap, gcap, device=local_rank)
            self.eng = self.slab.engine
            self.eng.set_colliders(scenes.engine_colliders(sc))
            self.eng.set_particles_ids(sc["pos"], sc["vel"], sc["ids"])
        else:
            self.eng = scenes.make_engine(sc, device=local_rank)
            self.eng.set_particles(sc["pos"], sc["vel"])
        self.n = self.n_global // world  # nominal particles per GPU (the slabs hold about this many each)
        self.cells = sc["grid"].total // world
        self.z_bounds = list(sc["z_bounds"]) if world > 1 else None
        self.rebalances = 0
        self.rebalance_ms = []
        # (sub-step, global plane histogram) of the last re-plan look; the initial plan IS a look at sub-step 0
        self.last_look = (0, __import__("numpy").asarray(sc["hist"], dtype="float64")) if world > 1 else None
        self.caps = [bb.slab_capacity(sc["hist"], sc["z_bounds"], r, slack=2.0)[0] for r in range(world)] if world > 1 else None

    def rebalance(self, every=100, gain=0.02):
        """Slab group: re-plan the z cuts from the per-plane histogram (bbx_plane_counts summed over the ranks ->
        bbx_slab_plan) and move there in neighbour-only steps (bbx_slab_plan_step, bbx_rebalance).  Host-synchronous;
        returns the wall-clock ms it took on this rank.  Policy: look at most once per `every` sub-steps (a look costs
        ~0.3 ms, a move ~2 ms: one to two sub-steps); plan on the histogram EXTRAPOLATED to the middle of the next
        interval from the drift since the previous look (a dam break pushes ~5 % of the mass across a cut per 100
        sub-steps); move only when the fullest slab of the new plan is at least `gain` x the mean lighter than the
        fullest slab of the current one.  Every decision is a function of the GLOBAL histogram and of the capacities
        every rank can compute, so all ranks take the same one (bbx_rebalance is collective); a step that would bring a
        slab within 10 % of its max_particles is not taken."""
        if self.world < 2:
            return 0.0
        import numpy as np
        import torch
        import torch.distributed as dist
        bb = self.bb
        t0 = time.perf_counter()
        now = int(self.eng.stats().substeps)
        if self.last_look is not None and now - self.last_look[0] < every:
            return 0.0
        hist = torch.from_numpy(self.eng.plane_counts()).cuda()
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
        hist = hist.cpu().numpy().astype(np.float64)
        plan_on = hist
        if self.last_look is not None and now > self.last_look[0]:
            drift = (hist - self.last_look[1]) / (now - self.last_look[0])   # particles per plane per sub-step
            plan_on = np.maximum(hist + drift * (every / 2), 0.0)
        self.last_look = (now, hist)
        target = [int(z) for z in bb.plan_slabs(np.rint(plan_on).astype(np.int64), self.world)]
        owned = lambda zb, h: [float(h[zb[r]:zb[r + 1]].sum()) for r in range(self.world)]
        if max(owned(self.z_bounds, plan_on)) - max(owned(target, plan_on)) < gain * plan_on.sum() / self.world:
            target = self.z_bounds
        moved = False
        for _ in range(8):
            if list(target) == self.z_bounds:
                break
            step, _ = bb.plan_step(self.z_bounds, target)
            step = [int(z) for z in step]
            if step == self.z_bounds:
                break
            if any(max(a, b) > 0.9 * c for a, b, c in zip(owned(step, hist), owned(step, plan_on), self.caps)):
                break   # (same verdict on every rank)
            self.eng.rebalance(step)
            self.z_bounds = step
            moved = True
        if moved:
            self.eng.synchronize()
            self.rebalances += 1
        ms = (time.perf_counter() - t0) * 1e3
        self.rebalance_ms.append(round(ms, 3))
        return ms

    def reset(self):
        """Single domain: back to the initial block.  (Slab groups keep stepping from where they are: a rank only holds
        its share of the initial block under the INITIAL cuts, and walking the cuts back there in neighbour-only steps can
        overfill a slab on the way.)"""
        if self.world == 1:
            self.eng.set_particles(self.sc["pos"], self.sc["vel"])

    def close(self):
        if self.eng is not None:
            self.eng.close()
            self.eng = None


def rank_max(values, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def barrier(world):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed_blocks(job, solver, dt, steps, repeats, local_rank, rebalance=True):
    """`repeats` back-to-back blocks of exactly `steps` sub-steps, each bracketed by two CUDA events on the engine's stream
    (bbx_step_many_timed) and a barrier + synchronize on both sides; per block the max over ranks.  Returns the per-block
    ms per sub-step, the wall-clock ms per sub-step of the median block, the launches per block and the clock record
    sampled DURING the blocks."""
    eng, world = job.eng, job.world
    sampler = ClockSampler(local_rank)
    barrier(world)
    sampler.start()
    blocks, walls = [], []
    l0 = eng.launches
    for _ in range(repeats):
        barrier(world)
        t0 = time.perf_counter()
        # slab groups re-plan their z cuts at the start of every block; that time (host wall clock, synchronous) is part of
        # the block: it is work the run needs in order to keep stepping at this rate
        reb = job.rebalance() if (world > 1 and rebalance) else 0.0
        ms = eng.step_many_timed(dt, steps, solver) + reb
        barrier(world)
        wall = (time.perf_counter() - t0) * 1e3
        ms, wall = rank_max([ms, wall], world)
        blocks.append(ms / steps)
        walls.append(wall / steps)
    clocks = sampler.stop()
    launches = (eng.launches - l0) // repeats
    order = sorted(range(repeats), key=lambda k: blocks[k])
    med = order[repeats // 2]
    return blocks, blocks[med], walls[med], launches, clocks


def phase_breakdown(job, solver, dt, steps, phase_ids):
    """Per-phase device time of `steps` further sub-steps (events between the kernels on the engine's stream)."""
    eng = job.eng
    eng.set_timing(True)
    eng.reset_kernel_time()
    eng.step_many(dt, steps, solver)
    eng.synchronize()
    phase = {name: eng.kernel_time(pid) for name, pid in phase_ids.items()}
    gap_ms, _ = eng.kernel_time(7)  # device idle time between sub-steps (launch gaps)
    eng.set_timing(False)
    return phase, gap_ms


def parity_leg(job, dt):
    """The parity gate on the state the timed region left behind (tests/parity_gate.py: the oracle as CHECKER, after the
    clock has stopped): one traced sub-step from identical inputs, cell order and neighbour lists bit-exact, fields
    within the FP32-vs-FP64 tolerances.  N > 1: every rank ships its owned rows (ids, FP32 state, cell order, one 64-bit
    checksum per neighbour-list row) to rank 0, which runs the oracle on the whole scene."""
    import numpy as np
    import torch.distributed as dist
    import parity_gate as pg
    import scenes
    bb, eng, world, rank, sc = job.bb, job.eng, job.world, job.rank, job.sc
    t0 = time.perf_counter()
    ext = float(np.max(np.asarray(sc["domain_max"]) - np.asarray(sc["domain_min"])))
    if world == 1:
        orc = scenes.make_oracle(sc)
        orc.set_particles(sc["pos"].astype(np.float64), sc["vel"].astype(np.float64))
        substep = eng.stats().substeps
        pg.sync_oracle_from_engine(eng, orc)
        r, _ = pg.gate_substep(eng, orc, dt, ext)
        r.update(substep=int(substep), particles=int(eng.n), oracle="oracle/bbx_oracle.c (FP64, pinned bit-exact to the unmodified reference)",
                 seconds=time.perf_counter() - t0, tolerances=pg.TOL)
        return r
    # ---- slabs: gather the inputs, trace on rank 0, compare what every rank computed
    ids, pos = eng.download_owned(bb.POSITION, np.float32)
    _, vel = eng.download_owned(bb.VELOCITY, np.float32)
    cc, co = eng.export_cells()
    st = eng.stats()
    inputs = [None] * world
    dist.gather_object((ids, pos, vel, np.nonzero(cc)[0].astype(np.int32), cc[cc > 0], co, int(st.rebuild_flag), int(st.substeps)), inputs if rank == 0 else None, dst=0)
    eng.run_phase(bb.PHASE_GRID, dt)
    eng.run_phase(bb.PHASE_DENSITY, dt)
    ids2, rho = eng.download_owned(bb.DENSITY, np.float32)
    cnt, rows = eng.export_neighbors_owned()
    sums = pg.row_checksums(rows)
    del rows
    cc2, co2 = eng.export_cells()
    outputs = [None] * world
    dist.gather_object((ids2, rho, cnt, sums, np.nonzero(cc2)[0].astype(np.int32), cc2[cc2 > 0], co2, int(eng.stats().exact_passes)), outputs if rank == 0 else None, dst=0)
    if rank != 0:
        return None
    n, total = job.n_global, sc["grid"].total
    p64, v64 = np.zeros((n, 3)), np.zeros((n, 3))
    cell_count, order, flag = np.zeros(total, dtype=np.int32), [], 0
    for i_, p_, v_, ci, cv, o_, f_, sub in inputs:
        p64[i_] = p_; v64[i_] = v_
        cell_count[ci] = cv; order.append(o_); flag |= f_
    orc = scenes.make_oracle(sc)
    orc.set_particles(p64, v64)
    orc.set_chains(cell_count, np.concatenate(order))  # slabs own ascending, disjoint cell ranges: global order = concatenation
    orc.S.rebuild_flag = flag
    tr = orc.trace_pcisph(dt)
    rho_e, cnt_e, sum_e = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.uint64)
    cc_e, order_e, exact = np.zeros(total, dtype=np.int32), [], 0
    owned = np.zeros(n, dtype=np.int32)
    for i_, r_, c_, s_, ci, cv, o_, x_ in outputs:
        rho_e[i_] = r_; cnt_e[i_] = c_; sum_e[i_] = s_; owned[i_] += 1
        cc_e[ci] = cv; order_e.append(o_); exact += x_
    r = {"substep": int(inputs[0][7]), "particles": int(n), "ranks": world,
         "every_particle_owned_once": bool((owned == 1).all()),
         "cell_counts_bit_exact": bool(np.array_equal(cc_e, tr["cell_count"])),
         "cell_order_bit_exact": bool(np.array_equal(np.concatenate(order_e), tr["cell_order"])),
         "neighbor_counts_bit_exact": bool(np.array_equal(cnt_e, tr["nbr_count"])),
         "neighbor_lists_bit_exact": bool(np.array_equal(sum_e, pg.row_checksums(tr["nbr_ids"]))),
         "neighbor_lists_compared_as": "64-bit order-sensitive checksum per list row (rows gathered from the ranks)",
         "err_density": float(np.abs(rho_e.astype(np.float64) - tr["density"]).max() / pg.RHO0), "exact_passes": int(exact),
         "oracle": "oracle/bbx_oracle.c on the whole scene, rank 0", "seconds": time.perf_counter() - t0, "tolerances": {"rho": pg.TOL["rho"]}}
    r["lists_bit_exact"] = bool(r["every_particle_owned_once"] and r["cell_counts_bit_exact"] and r["cell_order_bit_exact"]
                                and r["neighbor_counts_bit_exact"] and r["neighbor_lists_bit_exact"])
    r["fields_within_tolerance"] = bool(r["err_density"] < pg.TOL["rho"])
    r["ok"] = bool(r["lists_bit_exact"] and r["fields_within_tolerance"])
    return r


def e2e_leg(job, solver, dt, steps, sph):
    """The same sub-step through the C ABI with HOST buffers: every step uploads positions + velocities from pinned host
    memory, steps, and reads positions + velocities (+ ids on slabs) back into pinned host memory."""
    import torch
    bb, eng, world, sc = job.bb, job.eng, job.world, job.sc
    lib = eng.lib
    pos32, vel32 = sc["pos"], sc["vel"]
    hcap = max(len(pos32), int(eng.cfg.max_particles))  # slabs gain particles through migration
    buf = {k: torch.zeros((hcap, 3), dtype=torch.float32).pin_memory() for k in ("ip", "iv", "op", "ov")}
    hid = torch.zeros(hcap, dtype=torch.int32).pin_memory()
    cnt = C.c_int()
    h2d, d2h = [0], [0]
    job.reset()  # single domain: restart from the initial block; slab groups go on from the state they are in
    if world == 1:
        buf["ip"][:len(pos32)] = torch.from_numpy(pos32); buf["iv"][:len(vel32)] = torch.from_numpy(vel32)
        cnt.value = len(pos32)
    else:
        rc = lib.bbx_download_state(eng.h, buf["ip"].data_ptr(), buf["iv"].data_ptr(), hid.data_ptr(), bb.F32, 1, C.byref(cnt))
        if rc:
            raise SystemExit("bbx error: " + lib.bbx_last_error().decode())

    def step():
        m = cnt.value
        if world == 1:
            rc = lib.bbx_overwrite_state(eng.h, buf["ip"].data_ptr(), buf["iv"].data_ptr(), bb.F32)
            rc = rc or (lib.bbx_step_sph if sph else lib.bbx_step_pcisph)(eng.h, dt)
            rc = rc or lib.bbx_download_state(eng.h, buf["op"].data_ptr(), buf["ov"].data_ptr(), None, bb.F32, 0, C.byref(cnt))
            h2d[0] = 24 * m; d2h[0] = 24 * m
        else:
            # slab engines: every rank hands over the particles it holds (rows in the engine's cell order, the order of the
            # previous download), steps, and reads its owned particles back with their global ids
            rc = lib.bbx_overwrite_owned(eng.h, buf["ip"].data_ptr(), buf["iv"].data_ptr(), bb.F32)
            rc = rc or lib.bbx_step_pcisph(eng.h, dt)
            rc = rc or lib.bbx_download_state(eng.h, buf["op"].data_ptr(), buf["ov"].data_ptr(), hid.data_ptr(), bb.F32, 1, C.byref(cnt))
            h2d[0] = 24 * m; d2h[0] = 28 * cnt.value
        if rc:
            raise SystemExit("bbx error: " + lib.bbx_last_error().decode())
        buf["ip"], buf["op"] = buf["op"], buf["ip"]  # next step's input is this step's output
        buf["iv"], buf["ov"] = buf["ov"], buf["iv"]

    for _ in range(3):
        step()
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    barrier(world)
    sec = rank_max([(time.perf_counter() - t0) / steps], world)[0]
    tot = rank_max([float(h2d[0]), float(d2h[0])], 1)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor(tot, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        tot = [float(x) for x in t]
    return {"value": job.n_global / sec, "unit": UNIT, "h2d_bytes_per_step": int(tot[0]), "d2h_bytes_per_step": int(tot[1]),
            "ms_per_step": sec * 1e3,
            "api": ("bbx_overwrite_state + bbx_step_pcisph + bbx_download_state(positions, velocities), pinned host buffers" if world == 1 else
                    "per rank: bbx_overwrite_owned(host) + bbx_step_pcisph + bbx_download_state(positions, velocities, ids), pinned host buffers"),
            "steps": steps}


def extra_config(name, particles_total, workload, rank, world, local_rank, steps, warmup, peak):
    """One more BASELINE config on the same GPUs: ms per sub-step, updates/s, whole-step roofline fraction, clocks."""
    job = Job(particles_total, workload, rank, world, local_rank)
    try:
        job.eng.step_many(job.dt, warmup, job.bb.SOLVER_PCISPH)
        job.eng.synchronize()
        blocks, ms, wall, launches, clocks = timed_blocks(job, job.bb.SOLVER_PCISPH, job.dt, steps, 3, local_rank)
        st = job.eng.stats()
        gbs = (BYTES_PER_UPDATE * job.n + 8 * job.cells) / (ms * 1e-3) / 1e9
        return {"workload": name, "particles": job.n_global, "n_gpus": world, "ms_per_step": ms, "ms_per_step_blocks": blocks, "steps": steps, "warmup": warmup,
                "value": job.n_global / (ms * 1e-3), "unit": UNIT, "whole_step_frac": gbs / peak, "whole_step_gbs_per_gpu": gbs,
                "wall_ms_per_step": wall, "clocks": clocks, "nan_count": int(st.nan_count), "halo_p2p": bool(job.eng.p2p) if world > 1 else None}
    finally:
        job.close()


class Deadline:
    """Safety net of the bbx arm: the headline measurement is taken first and the legs behind it (parity gates, developed
    flow, e2e, extra configs, CPU baseline) fill further keys of the SAME line.  If the run is still going `seconds` after it
    started -- a leg slower than planned on this box, or one rank stuck in a collective another rank left with an error --
    rank 0 prints the line as it stands (missing keys null, "truncated" says why) and every rank exits 0, instead of the
    whole job dying at the driver's limit with nothing printed."""

    def __init__(self, seconds, rank):
        import threading
        self.rank, self.line, self.lock, self.done = rank, None, threading.Lock(), False
        self.t0 = time.perf_counter()
        self.timer = threading.Timer(seconds, self._fire, args=(f"deadline of {seconds:.0f} s reached",))
        self.timer.daemon = True
        if seconds > 0:
            self.timer.start()

    def _fire(self, why):
        with self.lock:
            if self.done:
                return
            self.done = True
            if self.rank == 0 and self.line is not None:
                self.line["truncated"] = why
                print(json.dumps(self.line), flush=True)
        os._exit(0 if (self.rank != 0 or self.line is not None) else 3)

    def fail(self, why):
        """an error on this rank after the headline was measured: same exit as the deadline"""
        self._fire(why)

    def finish(self):
        with self.lock:
            self.done = True
        self.timer.cancel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--repeats", type=int, default=5, help="timed blocks of --steps sub-steps; the median block is reported (BASELINE.md 3)")
    ap.add_argument("--impl", default="bbx", choices=["bbx", "reference"])
    ap.add_argument("--particles", type=float, default=1.0e6, help="particles per GPU (weak scaling)")
    ap.add_argument("--workload", default="dam", choices=["dam", "sdf"],
                    help="dam: PCISPH dam break in a box (configs 2, 4, 5); sdf: same plus a baked-SDF torus collider (config 3)")
    ap.add_argument("--solver", default="pcisph", choices=["pcisph", "sph"],
                    help="sph: the SphSolver3 step (BASELINE configs[0]; fixed dt 1.44e-4 = 0.4 h / c_s), single GPU line for the record")
    ap.add_argument("--ref-particles", type=float, default=1.0e6, help="reference arm: particles of its dam-break scene (default: the bench workload itself)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the additional BASELINE configs (8 M SDF at N = 1, 32 M at N = 2 / 4, 100 M at N = 8)")
    ap.add_argument("--developed-substeps", type=int, default=400, help="second timing after this many sub-steps (splash developed); 0 = skip")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--deadline", type=float, default=780.0, help="seconds after which the line is printed as it stands and every rank exits 0 (0 = off)")
    ap.add_argument("--no-rebalance", action="store_true", help="N > 1: keep the static slab plan of the initial distribution (default: re-plan the z cuts at the start of every timed block)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)
    guard = Deadline(args.deadline, rank)   # (counts from here: imports, scene generation and engine set-up included)

    import torch
    import torch.distributed as dist
    import bubbles_b200 as bb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: bubbles_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    job = Job(args.particles * world, args.workload, rank, world, local_rank)
    eng, n, n_global, cells = job.eng, job.n, job.n_global, job.cells
    dt = job.dt
    phase_ids, phase_bytes, bytes_per_update, solver = PHASE_IDS, PHASE_BYTES, BYTES_PER_UPDATE, bb.SOLVER_PCISPH
    if args.solver == "sph":
        phase_ids, phase_bytes, bytes_per_update, solver, dt = SPH_PHASE_IDS, SPH_PHASE_BYTES, SPH_BYTES_PER_UPDATE, bb.SOLVER_SPH, 1.44e-4
    state_bytes = n * (16 * 8 + 4 * 8 + 208)  # float4 arrays, scalars/indices, neighbour list
    peak, peak_src = peaks()

    # ---- device-resident throughput: W warm-up sub-steps, then `repeats` blocks of exactly K sub-steps -------------
    eng.step_many(dt, args.warmup, solver)
    eng.synchronize()
    blocks, ms_per_step, wall_ms, launches, clocks = timed_blocks(job, solver, dt, args.steps, args.repeats, local_rank, not args.no_rebalance)
    value = n_global / (ms_per_step * 1e-3)
    phase, gap_ms = phase_breakdown(job, solver, dt, args.steps, phase_ids)
    st = eng.stats()
    per_rank = None
    if world > 1:  # where the time of each rank goes (a phase ends with the wait for the neighbours' halo: skew shows up there)
        mine = {"rank": rank, "owned": int(st.particles), "ghosts": int(st.ghosts), "occupied_cells": int(st.occupied_cells),
                "phases_ms_per_step": {k: v[0] / args.steps for k, v in phase.items()}}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    if st.nan_count:
        raise SystemExit("non-finite positions during the timed region")
    stats = {"neighbor_overflow": st.neighbor_overflow, "clamped": st.clamped, "rebuild_flag": st.rebuild_flag, "occupied_cells": st.occupied_cells,
             "max_candidates": st.max_candidates, "exact_passes": st.exact_passes, "unstaged_tiles": st.unstaged_tiles, "substeps": st.substeps}
    p2p = bool(eng.p2p) if world > 1 else None

    # ---- the headline line (rank 0); the legs below fill its remaining keys --------------------------------------------
    line = None
    if rank == 0:
        dom = max(phase, key=lambda k: phase[k][0])
        dom_ms = phase[dom][0] / max(1, phase[dom][1])
        dom_bytes = phase_bytes[dom] * n + (8 * cells if dom == "grid" else 0)
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        traffic, limiter = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("particles") and dom in tj.get("dram_bytes_per_launch", {}):
                traffic = tj["dram_bytes_per_launch"][dom] * (n / tj["particles"])
            limiter = tj.get("limiters", {}).get(dom)
        step_gbs = (bytes_per_update * n + 8 * cells) / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": METRIC if args.solver == "pcisph" else "particle-updates/s (3D SPH sub-step, dam break)",
            "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_label(args.solver, args.workload, n_global, n, cells, world),
                       "particles_per_gpu": n, "particles": n_global, "cells": cells, "dt": dt,
                       "parallelism": f"slab{world}" if world > 1 else "single",
                       "halo": ("stores into the neighbours' ghost slots from inside the sweeps (CUDA IPC peer memory over NVLink) + NCCL for the global flags"
                                if p2p else "NCCL send/recv per phase") if world > 1 else None,
                       "l2_policy": f"working set {state_bytes / 1e6:.0f} MB per GPU > 126 MB L2 (no flush needed)",
                       "timing": f"{args.repeats} blocks of {args.steps} sub-steps after {args.warmup} warm-up, each block bracketed by two CUDA events on the engine's stream "
                                 "(launch gaps included), max over ranks per block; ms_per_step = the MEDIAN block",
                       "ms_per_step_blocks": blocks, "wall_ms_per_step": wall_ms, "host_numa_binding": numa,
                       "rebalance": (None if world == 1 else ("off (static plan)" if args.no_rebalance else
                                     "z cuts re-planned at the start of a timed block when >= 100 sub-steps have passed since the last look, on the drift-extrapolated "
                                     "plane histogram, moved when the fullest slab gets >= 2 % of the mean lighter (bbx_rebalance; host time inside the blocks)")),
                       "rebalances": job.rebalances, "rebalance_ms_rank0": list(job.rebalance_ms)},
            "clocks": clocks,
            "e2e": None,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_particle": phase_bytes[dom],
                         "whole_step": {"bytes_per_update": bytes_per_update, "achieved": step_gbs, "frac": step_gbs / peak},
                         "phases_ms_per_step": {k: v[0] / args.steps for k, v in phase.items()},
                         "gap_ms_per_step": gap_ms / args.steps,
                         "limiter": limiter},
            "stats": stats,
            "ranks": per_rank,
            "parity": None,
            "developed": None,
            "configs": {},
            "cpu_baseline": {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                             "sample": "only measured at N = 1" if world > 1 else "not measured"},
        }
        guard.line = line

    def put(key, val, sub=None):
        if line is not None:
            with guard.lock:   # (the deadline thread serialises the line under the same lock)
                (line if sub is None else line[sub])[key] = val

    try:
        # ---- end to end through the C ABI with host buffers (slab groups: from the state the timed region left) -------
        if world > 1:
            put("e2e", e2e_leg(job, solver, dt, args.e2e_steps, args.solver == "sph"))

        # ---- parity gate on the benched state (after the clock has stopped) -----------------------------------------
        if not args.no_parity and args.solver == "pcisph":
            put("parity", parity_leg(job, dt))

        # ---- the same measurement on a developed flow (splash, spray, compressed cells) -----------------------------
        if args.developed_substeps > 0 and args.solver == "pcisph":
            todo = args.developed_substeps - eng.stats().substeps
            while todo > 0:
                chunk = min(todo, 100)
                if world > 1 and not args.no_rebalance:
                    job.rebalance()
                eng.step_many(dt, chunk, solver)
                eng.synchronize()
                todo -= chunk
            dblocks, dms, dwall, _, dclocks = timed_blocks(job, solver, dt, args.steps, 3, local_rank, not args.no_rebalance)
            dphase, dgap = phase_breakdown(job, solver, dt, args.steps, phase_ids)
            dst = eng.stats()
            developed = {"after_substeps": int(args.developed_substeps), "ms_per_step": dms, "ms_per_step_blocks": dblocks, "value": n_global / (dms * 1e-3),
                         "phases_ms_per_step": {k: v[0] / args.steps for k, v in dphase.items()}, "clocks": dclocks,
                         "stats": {"exact_passes": dst.exact_passes, "unstaged_tiles": dst.unstaged_tiles, "max_candidates": dst.max_candidates,
                                   "occupied_cells": dst.occupied_cells, "clamped": dst.clamped, "nan_count": dst.nan_count}}
            put("developed", developed)
            if not args.no_parity:
                put("parity", parity_leg(job, dt), "developed")

        # ---- single domain: e2e restarts from the initial block, so it runs after the legs that need the flow ---------
        if world == 1:
            put("e2e", e2e_leg(job, solver, dt, args.e2e_steps, args.solver == "sph"))
        put("rebalances", job.rebalances, "config"); put("rebalance_ms_rank0", list(job.rebalance_ms), "config")
        job.close()

        # ---- the other BASELINE configs this GPU count can hold ------------------------------------------------------
        if not args.no_extra_configs and args.solver == "pcisph" and args.workload == "dam" and args.particles == 1.0e6:
            extra = {1: [("sdf8m", 8.0e6, "sdf")], 2: [("dam32m", 32.0e6, "dam")], 4: [("dam32m", 32.0e6, "dam")], 8: [("dam100m", 100.0e6, "dam")]}.get(world, [])
            for name, total, wl in extra:
                put(name, extra_config(name, total, wl, rank, world, local_rank, min(args.steps, 30), 10, peak), "configs")

        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            put("cpu_baseline", cpu_baseline(n))
            rg = reference_gpu(n)
            rg["note"] = ("extra key, not the reference arm: the unmodified reference in its native GPU mode on this box; the "
                          "`--impl reference` arm steps the CPU path (use_cpu = 1) and only touches the GPU while the reference's own setup code runs")
            put("reference_gpu", rg)
    except (Exception, SystemExit) as ex:  # a leg behind the headline failed on this rank: the line goes out as it stands, everybody leaves
        import traceback
        traceback.print_exc()
        if world > 1:
            time.sleep(2.0 if rank == 0 else 10.0)  # rank 0 prints first; the others give it that time, then leave too
        guard.fail(f"rank {rank}: {type(ex).__name__}: {str(ex)[-300:]}")
    guard.finish()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
