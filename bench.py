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
    est = n / 2.0e4 / max(1, min(cores, 8))  # s per sub-step, conservative
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
    sample = f"{n}-particle dam break (same constants as the GPU workload), {steps} sub-steps after {warm} warm-up, FP64"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": 1e3 * sec / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"PCISPH 3D dam break, bounded CPU sample of {n} particles, dt 7.2e-4, reference CPU path"},
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="bbx", choices=["bbx", "reference"])
    ap.add_argument("--particles", type=float, default=1.0e6, help="particles per GPU (weak scaling)")
    ap.add_argument("--workload", default="dam", choices=["dam", "sdf"],
                    help="dam: PCISPH dam break in a box (configs 2, 4, 5); sdf: same plus a baked-SDF torus collider (config 3)")
    ap.add_argument("--solver", default="pcisph", choices=["pcisph", "sph"],
                    help="sph: the SphSolver3 step (BASELINE configs[0]; fixed dt 1.44e-4 = 0.4 h / c_s), single GPU line for the record")
    ap.add_argument("--ref-particles", type=float, default=2.5e5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=10)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    import bubbles_b200 as bb
    import scenes

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: bubbles_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # weak scaling: ONE dam-break scene of (particles per GPU) x N particles, cut into N z-slabs of whole cell
    # planes balanced by particle count; ghost planes / migration / per-phase halos travel over NCCL.  Every rank
    # generates only its own share of the (deterministic) BCC block.
    sc = scenes.dam_break_scene_slab(args.particles * world, rank, world,
                                     obstacle=scenes.torus_obstacle if args.workload == "sdf" else None)
    n_global = sc["n_global"]
    pos32, vel32 = sc["pos"], sc["vel"]
    if world > 1:
        grid, zb, hist = sc["grid"], sc["z_bounds"], sc["hist"]
        cap, gcap = bb.slab_capacity(hist, zb, rank, slack=2.0)
        slab = bb.NcclSlab(grid, sc["spacing"], sc["scale"], zb, rank, world, broadcast_bytes, cap, gcap, device=local_rank)
        eng = slab.engine
        eng.set_colliders(scenes.engine_colliders(sc))
        eng.set_particles_ids(pos32, vel32, sc["ids"])
    else:
        eng = scenes.make_engine(sc, device=local_rank)
        eng.set_particles(pos32, vel32)
    n = n_global // world  # nominal particles per GPU (the slabs hold about this many each)
    dt = sc["dt"]
    phase_ids, phase_bytes, bytes_per_update, solver = PHASE_IDS, PHASE_BYTES, BYTES_PER_UPDATE, bb.SOLVER_PCISPH
    if args.solver == "sph":
        phase_ids, phase_bytes, bytes_per_update, solver, dt = SPH_PHASE_IDS, SPH_PHASE_BYTES, SPH_BYTES_PER_UPDATE, bb.SOLVER_SPH, 1.44e-4
    state_bytes = n * (16 * 8 + 4 * 8 + 208)  # float4 arrays, scalars/indices, neighbour list

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput -------------------------------------------------------------
    eng.step_many(dt, args.warmup, solver)
    eng.synchronize()
    eng.set_timing(True)
    eng.reset_kernel_time()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    l0 = eng.launches
    t0 = time.perf_counter()
    # CUDA events bracket the K sub-steps on the engine's own stream (bbx_advance-style timing is done
    # inside the library: per-phase events are recorded between the kernels, no sync until the end)
    eng.step_many(dt, args.steps, solver)
    eng.synchronize()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = eng.launches - l0
    phase = {}
    for name, pid in phase_ids.items():
        ms, k = eng.kernel_time(pid)
        phase[name] = (ms, k)
    gap_ms, _ = eng.kernel_time(7)  # device idle time between sub-steps (launch gaps), part of the step time
    ms_total = sum(v[0] for v in phase.values()) + gap_ms
    ms_per_step = ms_total / args.steps
    eng.set_timing(False)
    st = eng.stats()
    if st.nan_count:
        raise SystemExit("non-finite positions during the timed region")
    t = torch.tensor([ms_per_step, wall * 1e3 / args.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step, wall_ms = float(t[0]), float(t[1])
    value = n_global / (ms_per_step * 1e-3)

    # ---- end to end through the C ABI with host buffers ------------------------------------------
    hcap0 = max(len(pos32), int(eng.cfg.max_particles))
    hp = torch.zeros((hcap0, 3), dtype=torch.float32).pin_memory(); hp[:len(pos32)] = torch.from_numpy(pos32)
    hv = torch.zeros((hcap0, 3), dtype=torch.float32).pin_memory(); hv[:len(vel32)] = torch.from_numpy(vel32)
    op = torch.empty_like(hp).pin_memory()
    ov = torch.empty_like(hv).pin_memory()
    buf = {"ip": hp, "iv": hv, "op": op, "ov": ov}  # pinned input / output buffers, swapped after every step
    lib = eng.lib

    hcap = max(len(pos32), int(eng.cfg.max_particles))  # slabs gain particles through migration
    hid = torch.zeros(hcap, dtype=torch.int32).pin_memory()
    cnt = C.c_int()
    h2d = [0]
    d2h = [0]

    def e2e_step():
        if world == 1:
            rc = lib.bbx_overwrite_state(eng.h, buf["ip"].data_ptr(), buf["iv"].data_ptr(), bb.F32)
            rc |= (lib.bbx_step_sph if args.solver == "sph" else lib.bbx_step_pcisph)(eng.h, dt)
            rc |= lib.bbx_download(eng.h, bb.POSITION, buf["op"].data_ptr(), bb.F32)
            rc |= lib.bbx_download(eng.h, bb.VELOCITY, buf["ov"].data_ptr(), bb.F32)
            m = len(pos32)
            h2d[0] = 24 * m; d2h[0] = 24 * m
        else:
            # slab engines: every rank hands over the particles it holds (host buffers, global ids), steps, and
            # reads its owned particles back with their ids -- the per-rank share of what a host run loop does
            # (rows travel in the engine's cell order, the order of the previous download)
            m = cnt.value
            rc = lib.bbx_overwrite_owned(eng.h, buf["ip"].data_ptr(), buf["iv"].data_ptr(), bb.F32)
            rc |= lib.bbx_step_pcisph(eng.h, dt)
            rc |= lib.bbx_download_owned(eng.h, bb.POSITION, buf["op"].data_ptr(), bb.F32, hid.data_ptr(), C.byref(cnt))
            rc |= lib.bbx_download_owned(eng.h, bb.VELOCITY, buf["ov"].data_ptr(), bb.F32, None, None)
            h2d[0] = 24 * m; d2h[0] = 28 * m  # ids come back with the positions
        if rc:
            raise SystemExit("bbx error: " + lib.bbx_last_error().decode())
        # next step's input is this step's output: swap the pinned host buffers
        buf["ip"], buf["op"] = buf["op"], buf["ip"]
        buf["iv"], buf["ov"] = buf["ov"], buf["iv"]

    # restart from the initial block so that the e2e run simulates the same thing
    if world == 1:
        hp[:len(pos32)] = torch.from_numpy(pos32); hv[:len(vel32)] = torch.from_numpy(vel32)
        eng.set_particles(pos32, vel32)
    else:
        eng.set_particles_ids(pos32, vel32, sc["ids"])
        rc = lib.bbx_download_owned(eng.h, bb.POSITION, hp.data_ptr(), bb.F32, hid.data_ptr(), C.byref(cnt))
        rc |= lib.bbx_download_owned(eng.h, bb.VELOCITY, hv.data_ptr(), bb.F32, None, None)
        if rc:
            raise SystemExit("bbx error: " + lib.bbx_last_error().decode())
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n_global / float(t[0])

    if rank == 0:
        peak, peak_src = peaks()
        dom = max(phase, key=lambda k: phase[k][0])
        dom_ms = phase[dom][0] / max(1, phase[dom][1])
        cells = eng.grid.total // world
        dom_bytes = phase_bytes[dom] * n + (8 * cells if dom == "grid" else 0)
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("particles") and dom in tj.get("dram_bytes_per_launch", {}):
                traffic = tj["dram_bytes_per_launch"][dom] * (n / tj["particles"])
        step_gbs = (bytes_per_update * n + 8 * cells) / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": METRIC if args.solver == "pcisph" else "particle-updates/s (3D SPH sub-step, dam break)",
            "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("PCISPH 3D dam break + baked-SDF torus collider (BASELINE configs[2] stand-in), " if args.workload == "sdf" else "") +
                                   f"{'PCISPH' if args.solver == 'pcisph' else 'SPH'} 3D dam break, {n_global} particles = {n} per GPU (BASELINE configs[{'1' if args.solver == 'pcisph' else '0'}] per GPU), spacing 0.02, h = 1.8 s, " +
                                   (f"{cells} cells, fixed dt 7.2e-4, reference-compat (1 predict-correct iteration)" if args.solver == "pcisph" else
                                    f"{cells} cells, SphSolver3 step (BASELINE configs[0]), fixed dt 1.44e-4")
                                   + (f", {world} z-slabs with NCCL ghost-plane exchange and migration" if world > 1 else ""),
                       "particles_per_gpu": n, "particles": n_global, "cells": cells, "dt": dt,
                       "parallelism": f"slab{world}" if world > 1 else "single",
                       "halo": ("stores into the neighbours' ghost slots from inside the sweeps (CUDA IPC peer memory over NVLink) + NCCL for the grid phase"
                                if eng.p2p else "NCCL send/recv per phase") if world > 1 else None,
                       "l2_policy": f"working set {state_bytes / 1e6:.0f} MB per GPU > 126 MB L2 (no flush needed)",
                       "timing": "CUDA events on the engine stream between kernels, summed over phases, max over ranks",
                       "wall_ms_per_step": wall_ms},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d[0] * world, "d2h_bytes_per_step": d2h[0] * world,
                    "api": ("bbx_overwrite_state + bbx_step_pcisph + bbx_download(POSITION, VELOCITY), pinned host buffers" if world == 1 else
                            "per rank: bbx_overwrite_owned(host) + bbx_step_pcisph + bbx_download_owned(POSITION, VELOCITY, ids), pinned host buffers"),
                    "steps": args.e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_particle": phase_bytes[dom],
                         "whole_step": {"bytes_per_update": bytes_per_update, "achieved": step_gbs, "frac": step_gbs / peak},
                         "phases_ms_per_step": {k: v[0] / args.steps for k, v in phase.items()},
                         "gap_ms_per_step": gap_ms / args.steps},
            "stats": {"neighbor_overflow": st.neighbor_overflow, "clamped": st.clamped, "rebuild_flag": st.rebuild_flag,
                      "occupied_cells": st.occupied_cells, "max_candidates": st.max_candidates, "exact_passes": st.exact_passes, "unstaged_tiles": st.unstaged_tiles},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(n)
            eng.close(); eng = None  # free the device before the reference's own GPU build takes it
            line["reference_gpu"] = reference_gpu(n)
        else:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                                    "sample": "only measured at N = 1"}
        print(json.dumps(line), flush=True)
    if eng is not None:
        eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
