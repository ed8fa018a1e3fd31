"""CPU tests of the boundary: libbbx.so loads and exports every symbol include/bbx.h declares, the
host-only entry points work without a GPU, and the compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import bubbles_b200 as bb
from bubbles_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "bbx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bbx_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bbx.h but not exported by libbbx.so"
        assert n in L.SYMBOLS, f"{n} has no ctypes signature in bubbles_b200/_lib.py"
    assert lib.bbx_version() == 1


def test_struct_layout_matches_header():
    lib = L.load()
    cfg = L.Config()
    assert lib.bbx_config_default(C.byref(cfg), 1) == 0
    assert cfg.struct_size == C.sizeof(L.Config)       # the library wrote its own sizeof(bbx_config)
    assert cfg.viscosity == 0.04 and cfg.drag == 0.0001 and cfg.eos_exponent == 7.0
    assert cfg.sound_speed == 100.0 and cfg.pseudo_viscosity == 10.0 and cfg.pcisph_max_iterations == 5
    assert cfg.gravity[1] == float(np.float32(-9.8))    # vec3f(0.f, -9.8f, 0.f), sph_solver3.cpp:146
    assert cfg.restitution == 0.6 and cfg.time_step_limit_scale == 5.0


def test_grid_for_domain_matches_reference_facts():
    rows = np.load(os.path.join(ROOT, "tests", "golden", "grid_facts.npy"))
    for r in rows:
        g = bb.UtilBuildGridForDomain(r[0:3], r[3:6], r[6], r[7])
        assert np.array_equal(np.array(g.min[:]), r[8:11]) and np.array_equal(np.array(g.cell_len[:]), r[11:14])
        assert list(g.n) == [int(x) for x in r[14:17]] and g.total == int(np.prod(r[14:17]))
    g = bb.MakeGrid((10, 10, 10), (-1, -1, -1), (1, 1, 1))   # src/tests/test_grid.cpp:908-1007
    assert list(g.n) == [10, 10, 10] and abs(g.cell_len[0] - 0.2) < 1e-15


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    g = bb.UtilBuildGridForDomain((-0.3,) * 3, (0.3,) * 3, 0.02, 1.8)
    with pytest.raises(bb.BbxError) as ei:
        bb.Engine(g, 0.02, 1.8, 100)
    assert ei.value.code == L.ERR_NO_DEVICE


def test_product_does_not_touch_the_oracle():
    """The product path must not import, link or load anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "bubbles_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"


def test_python_emitter_glibc_mode_reproduces_the_reference_emitter():
    """bubbles_b200.emitter.VolumeParticleEmitter3(rng="glibc") against the particles the reference's emitter produced
    (tests/golden/probe_trace.npz: emit_box at (0.1, -0.1, 0.1), 0.2 x 0.3 x 0.2, jitter 0.001, srand(1))."""
    import numpy as np
    from bubbles_b200 import emitter
    g = np.load(os.path.join(ROOT, "tests", "golden", "probe_trace.npz"))
    c, size = np.array([0.1, -0.1, 0.1]), np.array([0.2, 0.3, 0.2])
    b = emitter.ParticleSetBuilder3()
    em = emitter.VolumeParticleEmitter3(emitter.box_inside_reference(c, size), c - size / 2, c + size / 2, 0.02, (0, -1, 0),
                                        jitter=0.001, seed=1, rng="glibc")
    em.Emit(b)
    assert b.GetParticleCount() == len(g["s0_pos"])
    assert np.array_equal(b.positions, g["s0_pos"]) and np.array_equal(b.velocities, g["s0_vel"])


def test_reference_side_binding_compiles_against_the_reference_headers():
    """bubbles_b200/host/reference_binding/pcisph_solver3_bbx.cpp (PciSphSolver3::Setup / SetColliders /
    Advance re-routed through the C ABI) is compiled by oracle/build_ref.sh against the unmodified reference headers and
    include/bbx.h; the object must exist, be newer than both sources and define the reference's member functions."""
    import subprocess
    src = os.path.join(ROOT, "bubbles_b200", "host", "reference_binding", "pcisph_solver3_bbx.cpp")
    obj = os.path.join(ROOT, "oracle", "_ref", "obj_bbx", "pcisph_solver3_bbx.o")
    if not os.path.isdir("/root/reference/src") and not os.path.exists(obj):
        pytest.skip("reference sources not present and no prebuilt object")
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh")], cwd=ROOT)
        assert os.path.getmtime(obj) >= os.path.getmtime(src) and os.path.getmtime(obj) >= os.path.getmtime(os.path.join(ROOT, "include", "bbx.h"))
    syms = subprocess.run(["nm", "-C", obj], capture_output=True, text=True).stdout
    for want in ("PciSphSolver3::Setup(", "PciSphSolver3::SetColliders(", "PciSphSolver3::Advance(", "PciSphSolver3::SetViscosityCoefficient("):
        assert any(want in l and " T " in l for l in syms.splitlines()), want
    for used in ("bbx_create", "bbx_set_particles", "bbx_set_colliders", "bbx_advance", "bbx_download", "bbx_update_collider"):
        assert any(l.strip().endswith(used) and " U " in l for l in syms.splitlines()), used
