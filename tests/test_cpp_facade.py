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
