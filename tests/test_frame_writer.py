"""SURVEY.md 8(f) row 1 -- frame files: the facade's SerializerSaveSphDataSet3 (bubbles_b200/host/bubbles_api.h)
against the reference's OWN serializer (src/third/serializer.cpp), through the unmodified-reference harness:
  * writer: the same particle state written by the reference (SaveSphParticleSet, :884-921) and by the facade
    must give byte-identical files, for every field combination the facade supports (p, pv, pvd, pvdm);
  * reader: the file the facade wrote is loaded by the reference's reader (SerializerLoadParticles3, :444-559 --
    what `bbtool view / pbr / surface` call) and must give the count, the format and the "%g" values back.
CPU only (the facade's particle set is a host mirror; no engine is involved)."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import __graft_entry__ as G
from oracle import oracle as O
import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "bubbles_b200", "lib", "frame_tool")
P, V, D, M = 0x01, 0x02, 0x04, 0x20

pytestmark = pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref/bbref not built (reference sources absent)")


def _tool():
    if not os.path.exists(TOOL):
        G.build()
    return TOOL


def _reference_state(tmp_path):
    """Probe scene advanced 3 sub-steps by the reference; returns the harness job prefix + dumped arrays."""
    sc = scenes.probe_scene()
    O.write_particles(str(tmp_path / "p.bin"), sc["pos"], sc["vel"])
    half = sc["domain_max"]
    I = O.mat_str(np.eye(4))
    job = ["threads 2", f"spacing {sc['spacing']}", f"scale {sc['scale']}",
           f"collider box {I} {float(2 * half[0])!r} {float(2 * half[1])!r} {float(2 * half[2])!r} 1 0", "domain_from_collider 0",
           f"particles {tmp_path}/p.bin", "setup", f"step {sc['dt']} 3", f"dump {tmp_path}/s_"]
    return sc, job


@pytest.mark.parametrize("flags", [P, P | V, P | V | D, P | V | D | M])
def test_writer_is_byte_identical_to_the_reference_and_its_reader_loads_it(tmp_path, flags):
    sc, job = _reference_state(tmp_path)
    ref_txt, our_txt = tmp_path / "ref.txt", tmp_path / "bbx.txt"
    out, _ = O.run_ref(job + [f"save_frame {ref_txt} {flags}"], str(tmp_path))
    pos, vel, rho = (np.load(tmp_path / f"s_{k}.npy") for k in ("pos", "vel", "density"))
    n = len(pos)
    # the mass the reference wrote (ParticleSet3::GetMass) is a Setup scalar: take it from the engine-independent
    # host arithmetic the facade uses (bbx_get_mass needs an engine; the oracle restates ComputeMass)
    orc = scenes.make_oracle(sc)
    mass = float(orc.P.mass)
    with open(tmp_path / "state.bin", "wb") as f:
        f.write(struct.pack("<qdd", n, float(sc["spacing"]), mass))
        f.write(np.ascontiguousarray(pos, np.float64).tobytes())
        f.write(np.ascontiguousarray(vel, np.float64).tobytes())
        f.write(np.ascontiguousarray(rho, np.float64).tobytes())
    r = subprocess.run([_tool(), str(tmp_path / "state.bin"), str(our_txt), str(flags)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(our_txt, "rb").read() == open(ref_txt, "rb").read()
    # the reference's reader on OUR file
    out, _ = O.run_ref([f"load_frame {our_txt} {tmp_path}/l_"], str(tmp_path))
    m = re.search(r"load_frame count=(\d+) flags=(\d+)", out)
    assert m and int(m.group(1)) == n and int(m.group(2)) == flags
    g = lambda a: np.array([float("%g" % x) for x in np.ravel(a)]).reshape(np.shape(a))  # what "%g" keeps
    # (the reference parses decimals with its own digit loop, ParseFloat: equal to strtod up to the last bits)
    same = lambda a, b: np.allclose(a, b, rtol=1e-14, atol=0.0)
    assert same(np.load(tmp_path / "l_pos.npy"), g(pos))
    if flags & V:
        assert same(np.load(tmp_path / "l_vel.npy"), g(vel))
    if flags & D:
        assert same(np.load(tmp_path / "l_rho.npy"), g(rho))
    if flags & M:
        assert same(np.load(tmp_path / "l_mass.npy"), np.full(n, float("%g" % mass)))


@pytest.mark.parametrize("sim_flags", [0x01 | 0x02, 0x01 | 0x02 | 0x08])
def test_simulation_file_with_shape_blocks_is_byte_identical(tmp_path, sim_flags):
    """(with the b flag the reference's UtilSaveSimulation3 always passes the boundary vector it builds from the particle set's
    v0 values -- all zero for a set nobody classified -- so the column is there)  UtilSaveSimulation3 (src/core/util.h:296-328): shape blocks (Shape::BoxSerialize / SphereSerialize,
    box.cpp:37-57, sphere.cpp:11-31) of every collider but the last -- the domain -- then the particle block."""
    sc = scenes.probe_scene()
    O.write_particles(str(tmp_path / "p.bin"), sc["pos"], sc["vel"])
    I = O.mat_str(np.eye(4))
    box_t, sph_t = (-0.15, -0.25, -0.1), (0.1, -0.27, 0.1)
    job = ["threads 2", f"spacing {sc['spacing']}", f"scale {sc['scale']}",
           f"collider box {O.mat_str(O.translate(*box_t))} 0.1 0.125 0.15 0 0.1",
           f"collider sphere {O.mat_str(O.translate(*sph_t))} 0.08 0 0.2",
           f"collider box {I} 0.6 0.6 0.6 1 0", "domain_from_collider 2",
           f"particles {tmp_path}/p.bin", "setup", f"step {sc['dt']} 2", f"dump {tmp_path}/s_",
           f"save_sim {tmp_path}/ref.txt {sim_flags}"]
    O.run_ref(job, str(tmp_path))
    pos, vel, rho = (np.load(tmp_path / f"s_{k}.npy") for k in ("pos", "vel", "density"))
    with open(tmp_path / "state.bin", "wb") as f:
        f.write(struct.pack("<qdd", len(pos), float(sc["spacing"]), 0.0))
        for a in (pos, vel, rho):
            f.write(np.ascontiguousarray(a, np.float64).tobytes())
    args = [_tool(), str(tmp_path / "state.bin"), str(tmp_path / "bbx.txt"), str(sim_flags),
            "--box", *[repr(float(x)) for x in box_t], "0.1", "0.125", "0.15",
            "--sphere", *[repr(float(x)) for x in sph_t], "0.08",
            "--box", "0", "0", "0", "0.6", "0.6", "0.6"]
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ref, ours = open(tmp_path / "ref.txt", "rb").read(), open(tmp_path / "bbx.txt", "rb").read()
    assert ref.startswith(b"ShapeBegin\n\t\"Type\" box") and ref.count(b"ShapeBegin") == 2  # the container is not written
    assert ours == ref


def _load_with_facade(path, tmp_path, tag):
    out = tmp_path / f"{tag}.bin"
    r = subprocess.run([_tool(), "--load", str(path), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    n, flags = struct.unpack("<qq", raw[:16])
    a = np.frombuffer(raw[16:], dtype=np.float64)
    return n, flags, a[:3 * n].reshape(n, 3), a[3 * n:6 * n].reshape(n, 3), a[6 * n:7 * n], a[7 * n:8 * n]


@pytest.mark.parametrize("flags", [P, P | V | D, P | V | D | M])
def test_reader_loads_reference_frames_bit_identically_to_the_reference_reader(tmp_path, flags):
    """The facade's frame reader (SerializerLoadParticles3 / LoadSphDataSet3 / LoadPoints3, serializer.cpp:84-261,
    444-577, numbers parsed like ParseFloat, obj_loader.cpp:270-278) on files the REFERENCE wrote -- plain frames and a
    simulation file with shape blocks in front -- against the reference's own reader on the same files: count, format
    and every value bit for bit."""
    sc, job = _reference_state(tmp_path)
    box = f"collider box {O.mat_str(O.translate(-0.15, -0.25, -0.1))} 0.1 0.1 0.1 0 0.1"
    job = job[:3] + [box] + job[3:]            # an obstacle in front of the container (written by save_sim)
    job[job.index("domain_from_collider 0")] = "domain_from_collider 1"
    ref_txt, sim_txt = tmp_path / "ref.txt", tmp_path / "sim.txt"
    O.run_ref(job + [f"save_frame {ref_txt} {flags}", f"save_sim {sim_txt} {flags}",
                     f"load_frame {ref_txt} {tmp_path}/a_", f"load_frame {sim_txt} {tmp_path}/b_"], str(tmp_path))
    for path, pre in ((ref_txt, "a_"), (sim_txt, "b_")):
        n, fl, pos, vel, rho, mass = _load_with_facade(path, tmp_path, pre)
        assert fl == flags and n == len(np.load(tmp_path / f"{pre}pos.npy"))
        assert np.array_equal(pos, np.load(tmp_path / f"{pre}pos.npy"))
        assert np.array_equal(vel, np.load(tmp_path / f"{pre}vel.npy"))
        assert np.array_equal(rho, np.load(tmp_path / f"{pre}rho.npy"))
        assert np.array_equal(mass, np.load(tmp_path / f"{pre}mass.npy"))


B, N, L, O_ = 0x08, 0x10, 0x40, 0x100   # boundary, normal, layers, boundary-exclusive rule (src/third/serializer.h:7-17)


@pytest.mark.parametrize("flags", [P | B, P | V | D | M | B | N, P | N, P | B | O_, P | V | B | N | L | O_, P | B | N])
def test_writer_boundary_and_normal_columns_are_byte_identical(tmp_path, flags):
    """The b / n columns and the boundary-exclusive rule (PushParticleSetToFile, serializer.cpp:812-880) with a boundary layer
    and normals a classification pass would have left in the particle set (here: synthetic values), written by the reference
    (SerializerSaveSphDataSet3 with the vector UtilGetBoundaryState builds) and by the facade: byte-identical files."""
    sc, job = _reference_state(tmp_path)
    rng = np.random.default_rng(3)
    n = len(sc["pos"])
    v0 = np.where(rng.random(n) < 0.3, rng.integers(1, 4, n), 0).astype(np.float64)
    nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True); nrm[v0 == 0] = 0.0
    for name, arr in (("b.bin", v0), ("n.bin", nrm)):
        with open(tmp_path / name, "wb") as f:
            f.write(struct.pack("<q", n)); f.write(np.ascontiguousarray(arr, np.float64).tobytes())
    ref_txt, our_txt = tmp_path / "ref.txt", tmp_path / "bbx.txt"
    O.run_ref(job + [f"set_boundary {tmp_path}/b.bin {tmp_path}/n.bin", f"save_frame_b {ref_txt} {flags}"], str(tmp_path))
    pos, vel, rho = (np.load(tmp_path / f"s_{k}.npy") for k in ("pos", "vel", "density"))
    mass = float(scenes.make_oracle(sc).P.mass)
    with open(tmp_path / "state.bin", "wb") as f:
        f.write(struct.pack("<qdd", n, float(sc["spacing"]), mass))
        for a in (pos, vel, rho):
            f.write(np.ascontiguousarray(a, np.float64).tobytes())
    r = subprocess.run([_tool(), str(tmp_path / "state.bin"), str(our_txt), str(flags), "--boundary", str(tmp_path / "b.bin"), str(tmp_path / "n.bin")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ours, ref = open(our_txt, "rb").read(), open(ref_txt, "rb").read()
    assert ours == ref
    count = int(re.search(rb'"Count" (\d+)', ref).group(1))
    assert count == (int((v0 > 0).sum()) if flags & O_ else n) and 0 < (v0 > 0).sum() < n


def test_writer_without_a_boundary_vector_keeps_the_reference_quirk(tmp_path):
    """No vector + the b flag: the reference warns, writes "b" into the format string and leaves the column out; the plain
    save_frame command of the harness does exactly that, and so does the facade."""
    sc, job = _reference_state(tmp_path)
    ref_txt, our_txt = tmp_path / "ref.txt", tmp_path / "bbx.txt"
    O.run_ref(job + [f"save_frame {ref_txt} {P | B}"], str(tmp_path))
    pos, vel, rho = (np.load(tmp_path / f"s_{k}.npy") for k in ("pos", "vel", "density"))
    with open(tmp_path / "state.bin", "wb") as f:
        f.write(struct.pack("<qdd", len(pos), float(sc["spacing"]), 0.0))
        for a in (pos, vel, rho):
            f.write(np.ascontiguousarray(a, np.float64).tobytes())
    r = subprocess.run([_tool(), str(tmp_path / "state.bin"), str(our_txt), str(P | B)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(our_txt, "rb").read() == open(ref_txt, "rb").read() and b'"Format" pb' in open(ref_txt, "rb").read()
