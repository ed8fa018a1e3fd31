"""bench.py's control flow on the CPU (host logic only): a FAKE engine stands in for the CUDA one so that the order of the
legs, the keys of the JSON line, the deadline net and the slab re-plan policy can be checked without a GPU.  Nothing here
measures anything; the real arm still refuses to run without a device (test_bench_contract.py)."""
import importlib.util
import json
import os
import subprocess
import sys
import time
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = r'''
import json, os, sys, time, types
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
import bench

class Stats:
    particles = 1000; ghosts = 0; occupied_cells = 10; nan_count = 0; neighbor_overflow = 0; clamped = 0; rebuild_flag = 0
    max_candidates = 300; exact_passes = 5; unstaged_tiles = 1
    def __init__(self, k): self.substeps = k

class Eng:
    p2p = False
    def __init__(self): self.k = 0; self.launches = 0
    def step_many(self, dt, n, solver): self.k += n; self.launches += 13 * n
    def step_many_timed(self, dt, n, solver): self.step_many(dt, n, solver); return 0.8 * n
    def synchronize(self): pass
    def stats(self): return Stats(self.k)
    def set_timing(self, on): pass
    def reset_kernel_time(self): pass
    def kernel_time(self, pid): return (0.1 * (pid + 1), 1)

class Job:
    def __init__(self, *a): self.eng = Eng(); self.n = self.n_global = 1000; self.cells = 100; self.dt = 7.2e-4; self.world = 1; self.rank = 0; self.rebalances = 0; self.rebalance_ms = []
    def rebalance(self): return 0.0
    def reset(self): pass
    def close(self): pass

class Sampler:
    def __init__(self, i): pass
    def start(self): pass
    def stop(self): return {{"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 3, "source": "fake"}}

bench.Job, bench.ClockSampler = Job, Sampler
bench.bind_to_gpu_numa_node = lambda i: None
bench.rank_max = lambda values, world: [float(v) for v in values]
bench.barrier = lambda world: None
bench.parity_leg = lambda job, dt: {{"ok": True}}
def e2e(job, solver, dt, steps, sph):
    if {hang!r}: time.sleep(60)
    if {boom!r}: raise RuntimeError("boom in a leg")
    return {{"value": 1.0, "unit": bench.UNIT, "h2d_bytes_per_step": 24, "d2h_bytes_per_step": 24}}
bench.e2e_leg = e2e
bench.extra_config = lambda name, *a: {{"workload": name, "ms_per_step": 1.0}}
bench.cpu_baseline = lambda n: {{"value": 2.0, "unit": bench.UNIT, "cores": 1, "kind": "reference", "sample": "fake"}}
bench.reference_gpu = lambda n: {{"value": 3.0}}
sys.argv = ["bench.py", "--steps", "4", "--warmup", "3", "--deadline", {deadline!r}]
bench.main()
'''


def run(hang=False, boom=False, deadline="60"):
    r = subprocess.run([sys.executable, "-c", DRIVER.format(root=ROOT, hang=hang, boom=boom, deadline=deadline)], capture_output=True, text=True, timeout=300)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    return r, lines


def test_line_has_every_contract_key_and_the_legs_fill_it():
    r, lines = run()
    assert r.returncode == 0, r.stderr[-3000:]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "parity", "developed", "configs", "stats"):
        assert k in d, k
    assert d["steps"] == 4 and d["warmup"] == 3 and d["n_gpus"] == 1 and d["ms_per_step"] == pytest.approx(0.8)
    assert d["value"] == pytest.approx(1000 / 0.8e-3) and d["gpu_launches"] == 13 * 4
    assert d["e2e"]["h2d_bytes_per_step"] == 24 and d["parity"] == {"ok": True} and d["developed"]["parity"] == {"ok": True}
    assert d["configs"]["sdf8m"]["workload"] == "sdf8m" and d["cpu_baseline"]["value"] == 2.0 and "truncated" not in d
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])


def test_deadline_prints_the_headline_as_it_stands_and_exits_zero():
    r, lines = run(hang=True, deadline="5")
    assert r.returncode == 0, r.stderr[-3000:]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["value"] > 0 and d["e2e"] is None and "deadline" in d["truncated"] and d["parity"] == {"ok": True}


def test_a_failing_leg_does_not_cost_the_headline():
    r, lines = run(boom=True)
    assert r.returncode == 0, r.stderr[-3000:]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["value"] > 0 and d["e2e"] is None and "boom in a leg" in d["truncated"]


def test_slab_replan_policy_is_rank_independent_and_respects_capacities():
    """Job.rebalance on fake engines: the look happens at most every 100 sub-steps, the plan uses the drift-extrapolated
    histogram, small gains do not move the cuts, a step that would overfill a slab is not taken."""
    import bench
    import bubbles_b200 as bb
    job = bench.Job.__new__(bench.Job)
    moves = []

    class Eng:
        k = 0
        hist = None
        def stats(self): return types.SimpleNamespace(substeps=self.k)
        def plane_counts(self): return self.hist.copy()
        def rebalance(self, zb): moves.append(list(zb))
        def synchronize(self): pass

    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("host-logic test for the CPU suite")
    # stand-ins for the two torch calls of the look (one rank: the all-reduce is the identity)
    import torch.distributed as dist
    real_cuda, real_ar = torch.Tensor.cuda, dist.all_reduce
    torch.Tensor.cuda = lambda self, *a, **k: self
    dist.all_reduce = lambda t, op=None: None
    try:
        hist0 = np.zeros(40, dtype=np.int64); hist0[20:40] = 1000            # 20 000 particles in planes 20..39
        job.bb, job.world, job.rank, job.eng = bb, 2, 0, Eng()
        job.z_bounds, job.rebalances, job.rebalance_ms = [0, 30, 40], 0, []
        job.last_look = (0, hist0.astype(np.float64))
        job.caps = [20000 + 1000, 20000 + 1000]
        job.eng.hist = hist0
        job.eng.k = 50
        assert job.rebalance() == 0.0 and moves == []                          # too early to look
        job.eng.k = 100
        job.rebalance()
        assert moves == [] and job.z_bounds == [0, 30, 40]                     # balanced: nothing to gain
        h = hist0.copy(); h[20:24] += 700; h[36:40] -= 700                      # mass drifts towards low z
        job.eng.hist, job.eng.k = h, 200
        job.rebalance()
        assert len(moves) == 1 and moves[0][1] < 30 and job.z_bounds == moves[0]   # the cut follows (and leads) the drift
        job.caps = [100, 100]                                                   # nobody has room: no step is taken
        h2 = h.copy(); h2[20:24] += 2000; h2[36:40] -= 200
        job.eng.hist, job.eng.k = h2, 300
        before = list(job.z_bounds)
        job.rebalance()
        assert job.z_bounds == before and len(moves) == 1
    finally:
        torch.Tensor.cuda, dist.all_reduce = real_cuda, real_ar


# ------------------------------------------------------------------------------------------------ two ranks over gloo
DRIVER2 = r'''
import json, os, sys, time
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch
import torch.distributed as dist
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.Tensor.cuda = lambda self, *a, **k: self
_tensor = torch.tensor
torch.tensor = lambda *a, **k: _tensor(*a, **{{kk: vv for kk, vv in k.items() if kk != "device"}})
_init = dist.init_process_group
dist.init_process_group = lambda backend=None, **k: _init("gloo")          # the plumbing is the same, the wire is not NCCL
import bench
RANK = int(os.environ["RANK"])

class Stats:
    ghosts = 30; occupied_cells = 10; nan_count = 0; neighbor_overflow = 0; clamped = 0; rebuild_flag = 0
    max_candidates = 300; exact_passes = 5; unstaged_tiles = 1
    def __init__(self, k): self.substeps = k; self.particles = 1000 + 10 * RANK

class Eng:
    p2p = True
    def __init__(self): self.k = 0; self.launches = 0
    def step_many(self, dt, n, solver): self.k += n; self.launches += 15 * n
    def step_many_timed(self, dt, n, solver): self.step_many(dt, n, solver); return (0.9 + 0.1 * RANK) * n
    def synchronize(self): pass
    def stats(self): return Stats(self.k)
    def set_timing(self, on): pass
    def reset_kernel_time(self): pass
    def kernel_time(self, pid): return (0.1 * (pid + 1), 1)

class Job:
    def __init__(self, total, workload, rank, world, local_rank):
        self.eng = Eng(); self.n = 1000; self.n_global = 2000; self.cells = 100; self.dt = 7.2e-4; self.world = world; self.rank = rank
        self.rebalances = 0; self.rebalance_ms = []
    def rebalance(self): return 0.0
    def reset(self): pass
    def close(self): pass

class Sampler:
    def __init__(self, i): pass
    def start(self): pass
    def stop(self): return {{"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 3, "source": "fake"}}

bench.Job, bench.ClockSampler = Job, Sampler
bench.bind_to_gpu_numa_node = lambda i: None
bench.parity_leg = lambda job, dt: ({{"ok": True}} if job.rank == 0 else None)
def e2e(job, solver, dt, steps, sph):
    if {boom_rank!r} == job.rank: raise RuntimeError("boom on one rank")
    return {{"value": 1.0, "unit": bench.UNIT, "h2d_bytes_per_step": 48, "d2h_bytes_per_step": 56}}
bench.e2e_leg = e2e
bench.extra_config = lambda name, *a: {{"workload": name, "ms_per_step": 12.0}}
sys.argv = ["bench.py", "--gpus", "2", "--steps", "4", "--warmup", "3", "--deadline", {deadline!r}]
bench.main()
'''


def run2(boom_rank=-1, deadline="90"):
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", DRIVER2.format(root=ROOT, boom_rank=boom_rank, deadline=deadline)],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=280) for p in procs]
    return procs, outs


def test_two_ranks_over_gloo_print_one_line_with_the_max_over_ranks():
    procs, outs = run2()
    for p, (_, err) in zip(procs, outs):
        assert p.returncode == 0, err[-3000:]
    lines0 = [l for l in outs[0][0].splitlines() if l.startswith("{")]
    lines1 = [l for l in outs[1][0].splitlines() if l.startswith("{")]
    assert len(lines0) == 1 and lines1 == []                                     # rank 0 alone prints
    d = json.loads(lines0[0])
    assert d["n_gpus"] == 2 and d["ms_per_step"] == pytest.approx(1.0)              # rank 1 is the slower one: max over ranks
    assert d["value"] == pytest.approx(2000 / 1.0e-3) and d["config"]["parallelism"] == "slab2"
    assert [r["rank"] for r in d["ranks"]] == [0, 1] and d["ranks"][1]["owned"] == 1010
    assert d["e2e"]["d2h_bytes_per_step"] == 56 and d["parity"] == {"ok": True} and d["configs"]["dam32m"]["ms_per_step"] == 12.0
    assert d["cpu_baseline"]["value"] is None and "truncated" not in d


def test_a_leg_failing_on_one_rank_still_yields_the_headline_and_nobody_hangs():
    t0 = time.time() if False else None
    procs, outs = run2(boom_rank=1, deadline="40")
    lines0 = [l for l in outs[0][0].splitlines() if l.startswith("{")]
    assert procs[0].returncode == 0 and procs[1].returncode == 0, (outs[0][1][-2000:], outs[1][1][-2000:])
    assert len(lines0) == 1
    d = json.loads(lines0[0])
    assert d["value"] > 0 and "truncated" in d
