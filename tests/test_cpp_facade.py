"""The C++ host facade (bubbles_b200/host/bubbles_api.h) and its demo scene (the reference's
test_pcisph3_dam_break, src/tests/test_pcisph_extra.cpp:1102-1169): builds with plain g++ (host only), fails
loudly without a GPU, and on a GPU produces exactly what the ctypes path produces for the same scene, plus a
text frame in the format bbtool reads (src/third/serializer.cpp:884-921)."""
import os
import re
import subprocess

import numpy as np
import pytest

import __graft_entry__ as G
import bubbles_b200 as bb
import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "bubbles_b200", "lib", "dam_break_demo")
HEADER = os.path.join(ROOT, "bubbles_b200", "host", "bubbles_api.h")


def _demo():
    if not os.path.exists(DEMO):
        G.build()
    return DEMO


def test_facade_header_is_host_only_cxx():
    # no nvcc, no CUDA headers: a Bubbles scene script compiles against it with the host compiler alone
    src = '#include "%s"\nint main(){ bbx::PciSphSolver3 s; bbx::SphSolver3 t; (void)s; (void)t; return 0; }\n' % HEADER
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-x", "c++", "-"], input=src, text=True, capture_output=True)
    assert r.returncode == 0, r.stderr


def test_demo_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([_demo(), "--steps", "1"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr


def _python_twin(steps, dt):
    ds = float(np.float32(0.6))
    spacing, scale = float(np.float32(0.02)), float(np.float32(1.8))
    fl, fy, bl, by = 0.5 * ds, 0.9 * ds, 1.3 * ds, 1.2 * ds
    xof = (bl - fl) / 2.0 - spacing
    zof = (bl - fl) / 2.0 - spacing
    yof = (by - fy) / 2.0 - spacing
    sc = scenes.block_scene((bl, by, bl), (fl, fy, fl), (xof, -yof, zof), (0, -6, 0), spacing=spacing, scale=scale, jitter=0.0, dt=dt)
    eng = scenes.make_engine(sc)
    eng.set_particles(sc["pos"], sc["vel"])
    eng.step_many(dt, steps)
    return sc, eng


@pytest.mark.gpu
def test_demo_matches_ctypes_path_and_writes_bbtool_frames(tmp_path):
    steps, dt = 25, 7.2e-4
    dump = tmp_path / "state.bin"
    r = subprocess.run([_demo(), "--jitter", "0", "--steps", str(steps), "--dt", repr(dt), "--dump", str(dump), "--out", str(tmp_path)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(dump, dtype=np.float64)
    n = int(raw[0])
    pos = raw[1:1 + 3 * n].reshape(n, 3)
    vel = raw[1 + 3 * n:1 + 6 * n].reshape(n, 3)
    sc, eng = _python_twin(steps, dt)
    assert n == len(sc["pos"])
    assert np.array_equal(pos, eng.download(bb.POSITION))      # same library, same inputs: bit-identical
    assert np.array_equal(vel, eng.download(bb.VELOCITY))
    # frame 0 = the emitted block, frame 1 = after the sub-steps; format of SaveSphParticleSet
    for frame, ref in ((0, sc["pos"]), (1, pos)):
        lines = open(tmp_path / f"out_{frame}.txt").read().split("\n")
        assert lines[0] == "FluidBegin" and lines[1] == '\t"Type" particles'
        assert lines[2] == f'\t"Count" {n}' and lines[3] == '\t"Format" p'
        assert lines[4] == '\t"Spacing" %g' % float(np.float32(0.02)) and lines[5] == "\tDataBegin"
        assert lines[6 + n] == "\tDataEnd" and lines[7 + n] == "FluidEnd"
        for k in (0, n // 2, n - 1):
            assert lines[6 + k] == "\t\t%g %g %g" % tuple(ref[k])


@pytest.mark.gpu
def test_demo_advance_frames_and_sph():
    r = subprocess.run([_demo(), "--frames", "2"], capture_output=True, text=True)
    assert r.returncode == 0 and "===== OK" in r.stdout, r.stdout + r.stderr
    assert "non-finite 0" in r.stdout
    r = subprocess.run([_demo(), "--sph", "--steps", "10"], capture_output=True, text=True)
    assert r.returncode == 0 and "===== OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_demo_continuous_emission():
    """ContinuousParticleSetBuilder3 of the facade: particles appended between frames (bbx_append_particles after
    stepping) -- the particle count grows by K per frame and the run stays finite."""
    r = subprocess.run([_demo(), "--frames", "3", "--emit", "100"], capture_output=True, text=True)
    assert r.returncode == 0 and "===== OK" in r.stdout, r.stdout + r.stderr
    counts = [int(x) for x in __import__("re").findall(r"Particles (\d+)", r.stdout)]
    assert len(counts) == 3 and counts[1] == counts[0] + 100 and counts[2] == counts[0] + 200, r.stdout
    assert "non-finite 0" in r.stdout


@pytest.mark.gpu
def test_demo_map_grid_emit_runs_on_the_gpu():
    """The facade's MapGrid / MapGridEmit (src/core/grid.h:1288-1407) end to end on the device: the cells the block
    started in are refilled after every frame (per-cell test: bbx_query_cells, Commit: bbx_append_particles)."""
    r = subprocess.run([_demo(), "--frames", "3", "--map-emit"], capture_output=True, text=True)
    assert r.returncode == 0 and "===== OK" in r.stdout, r.stdout + r.stderr
    added = [int(x) for x in re.findall(r"map-emit added (\d+)", r.stdout)]
    counts = [int(x) for x in re.findall(r"Particles (\d+)", r.stdout)]
    assert len(added) == 2 and all(a > 0 for a in added), r.stdout
    assert counts[1] == counts[0] + added[0] and counts[2] == counts[1] + added[1], r.stdout
    assert "non-finite 0" in r.stdout


@pytest.mark.gpu
def test_demo_restarts_from_a_frame_file(tmp_path):
    """--load: a frame written by the facade (positions only, "%g") is read back by the facade's reader and stepped."""
    r = subprocess.run([_demo(), "--jitter", "0", "--steps", "5", "--out", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    n = int(re.search(r"particles (\d+),", r.stdout).group(1))
    r = subprocess.run([_demo(), "--load", str(tmp_path / "out_1.txt"), "--steps", "5"], capture_output=True, text=True)
    assert r.returncode == 0 and "===== OK" in r.stdout, r.stdout + r.stderr
    assert f"loaded {n} particles (format p)" in r.stdout and "non-finite 0" in r.stdout


def test_facade_emitter_reproduces_the_reference_emitter_bit_for_bit(tmp_path):
    """VolumeParticleEmitter3 of the facade (BCC lattice, jitter from libc rand(), shape test) against the particles the
    REFERENCE's emitter produced for the same box (tests/golden/probe_trace.npz, s0_*: emit_box at (0.1, -0.1, 0.1),
    0.2 x 0.3 x 0.2, spacing 0.02, v = (0, -1, 0), jitter 0.001, srand(1)).  CPU only."""
    G.build()
    tool = os.path.join(ROOT, "bubbles_b200", "lib", "frame_tool")
    out = tmp_path / "emit.bin"
    r = subprocess.run([tool, "--emit", str(out), "0.1", "-0.1", "0.1", "0.2", "0.3", "0.2", "0.02", "0", "-1", "0", "0.001", "1"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    n = int(np.frombuffer(raw[:8], dtype=np.int64)[0])
    a = np.frombuffer(raw[8:], dtype=np.float64)
    g = np.load(os.path.join(ROOT, "tests", "golden", "probe_trace.npz"))
    assert n == len(g["s0_pos"])
    assert np.array_equal(a[:3 * n].reshape(n, 3), g["s0_pos"])
    assert np.array_equal(a[3 * n:].reshape(n, 3), g["s0_vel"])


def test_facade_map_grid_emit_rule_matches_the_reference(tmp_path):
    """ContinuousParticleSetBuilder3::MapGridEmit of the facade (MapGridEmitCandidates: which template positions are
    re-emitted) on the states the REFERENCE went through (tests/golden/emit_run.npz: chains and positions before each
    of its two emissions) must pick exactly the particles the reference added, in its order.  CPU only."""
    G.build()
    tool = os.path.join(ROOT, "bubbles_b200", "lib", "frame_tool")
    g = np.load(os.path.join(ROOT, "tests", "golden", "emit_run.npz"))
    n0 = len(g["p_pos"])
    cc0, co0 = g["s0_cell_count"], g["s0_cell_order"]
    start0 = np.concatenate([[0], np.cumsum(cc0)])
    mapped = {int(c): g["p_pos"][co0[start0[c]:start0[c + 1]]] for c in np.nonzero(cc0)[0]}   # MapGrid right after setup

    def run(pos, cc, co):
        with open(tmp_path / "in.bin", "wb") as f:
            f.write(np.float64(0.02).tobytes())
            f.write(np.array([len(pos), len(cc), len(mapped)], dtype=np.int64).tobytes())
            f.write(np.ascontiguousarray(pos, np.float64).tobytes())
            f.write(np.ascontiguousarray(cc, np.int32).tobytes())
            f.write(np.ascontiguousarray(co, np.int32).tobytes())
            for c in sorted(mapped):
                f.write(np.array([c, len(mapped[c])], dtype=np.int64).tobytes())
                f.write(np.ascontiguousarray(mapped[c], np.float64).tobytes())
        r = subprocess.run([tool, "--mapemit", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        raw = open(tmp_path / "out.bin", "rb").read()
        k = int(np.frombuffer(raw[:8], dtype=np.int64)[0])
        return np.frombuffer(raw[8:], dtype=np.float64).reshape(k, 3)

    add1 = run(g["s40_pos"], g["s40_cell_count"], g["s40_cell_order"])
    assert len(add1) == int(g["added"][0]) and np.array_equal(add1, g["e1_pos"][n0:])
    # second emission: the state right before it is not in the golden, but the oracle (pinned on this very run) is
    from oracle import oracle as O
    orc = O.Oracle(0.02, 1.8, (-0.3, -0.3, -0.3), (0.3, 0.3, 0.3), [O.make_collider("box", size=(0.6, 0.6, 0.6), reverse=True)])
    orc.set_particles(g["e1_pos"], g["e1_vel"])
    orc.set_chains(g["e1_cell_count"], g["e1_cell_order"])
    for _ in range(30):
        orc.substep_pcisph(7e-4)
    add2 = run(orc.a["pos"], orc.arr("cell_count"), orc.arr("cell_order"))
    assert len(add2) == int(g["added"][1]) and np.array_equal(add2, g["e2_pos"][len(g["e1_pos"]):])
