"""Host-side particle emission (setup time, not on the step path).

VolumeParticleEmitter3 / VolumeParticleEmitterSet3 / ParticleSetBuilder3 of the reference
(src/core/emitter.cpp:242-330, src/generator/bcclattice.cpp:5-36, src/core/particle.h:611-710):
a body-centred-cubic lattice over the emitter bounds, each point jittered by
0.5 * jitter * spacing * (random unit vector) and kept when it lies inside the shape.
The reference draws the jitter from libc rand(); by default a numpy Generator is used here (seeded), so
emitted sets are reproducible across platforms but not bit-identical to a glibc run -- parity
tests therefore feed both sides the same explicit positions.  `rng="glibc"` reproduces the reference's
emitter bit for bit on a glibc host (libc srand / rand, the draw order of its build, the box test with the
reference's Inside slack: box_inside_reference), pinned by tests/test_abi.py against tests/golden/probe_trace.npz.
"""
import ctypes
import ctypes.util

import numpy as np


def bcc_lattice(bounds_min, bounds_max, spacing):
    """Points of BccLatticePointGenerator::ForEach in its iteration order (z layers, y, x)."""
    lo = np.asarray(bounds_min, dtype=np.float64)
    hi = np.asarray(bounds_max, dtype=np.float64)
    half = spacing / 2
    ext = np.abs(hi - lo)
    out = []
    k = 0
    shifted = False
    while k * half <= ext[2]:
        off = half if shifted else 0.0
        ny = 0
        while ny * spacing + off <= ext[1]:
            ny += 1
        nx = 0
        while nx * spacing + off <= ext[0]:
            nx += 1
        if nx and ny:
            x = np.arange(nx) * spacing + off + lo[0]
            y = np.arange(ny) * spacing + off + lo[1]
            yy, xx = np.meshgrid(y, x, indexing="ij")
            zz = np.full_like(xx, k * half + lo[2])
            out.append(np.stack([xx, yy, zz], axis=-1).reshape(-1, 3))
        shifted = not shifted
        k += 1
    return np.concatenate(out, axis=0) if out else np.zeros((0, 3))


class ParticleSetBuilder3:
    def __init__(self):
        self.positions = np.zeros((0, 3))
        self.velocities = np.zeros((0, 3))

    def AddParticles(self, pos, vel):
        self.positions = np.concatenate([self.positions, np.asarray(pos, dtype=np.float64).reshape(-1, 3)])
        self.velocities = np.concatenate([self.velocities, np.asarray(vel, dtype=np.float64).reshape(-1, 3)])

    def AddParticle(self, pos, vel=(0, 0, 0)):
        self.AddParticles([pos], [vel])
        return 1

    def SetVelocityForAll(self, vel):
        self.velocities[:] = vel

    def GetParticleCount(self):
        return len(self.positions)


class VolumeParticleEmitter3:
    """One-shot volume emitter. `inside(points) -> bool[n]` is the shape's SignedDistance(p) <= 0 test."""

    def __init__(self, inside, bounds_min, bounds_max, spacing, init_vel=(0, 0, 0), jitter=0.0, seed=1,
                 max_particles=None, rng="numpy"):
        self.inside = inside
        self.rng_kind = rng
        self.seed = seed
        self.bounds = (np.asarray(bounds_min, float), np.asarray(bounds_max, float))
        self.spacing = spacing
        self.init_vel = np.asarray(init_vel, float)
        self.jitter = float(np.clip(jitter, 0.0, 1.0))
        self.rng = np.random.default_rng(seed)
        self.max_particles = max_particles

    def SetJitter(self, jitter):
        self.jitter = float(np.clip(jitter, 0.0, 1.0))

    def Emit(self, builder):
        pts = bcc_lattice(self.bounds[0], self.bounds[1], self.spacing)
        if self.rng_kind == "glibc" and len(pts):
            # emitter.cpp:242-330: one (u0, u1) pair per lattice point, drawn even when jitter = 0; the reference builds
            # vec2f u(rand_float(), rand_float()) and its compiler evaluates the arguments right to left: u[1] first
            libc = ctypes.CDLL(ctypes.util.find_library("c"))
            libc.srand(ctypes.c_uint(self.seed))
            raw = np.array([libc.rand() for _ in range(2 * len(pts))], dtype=np.float64).reshape(-1, 2)
            u = (raw.astype(np.float32) / np.float32(2147483648.0)).astype(np.float64)  # rand() / (RAND_MAX + 1.f) in float
            u1, u0 = u[:, 0], u[:, 1]
            usqrt = 2 * np.sqrt(u1 * (1 - u1))
            utheta = 2 * np.pi * u0
            d = np.stack([np.cos(utheta) * usqrt, np.sin(utheta) * usqrt, 1 - 2 * u1], axis=-1)
            pts = pts + (0.5 * self.jitter * self.spacing) * d
        elif self.jitter > 0 and len(pts):
            # SampleSphere(u): uniform direction on the unit sphere (geometry.h)
            u = self.rng.random((len(pts), 2))
            z = 1 - 2 * u[:, 0]
            r = np.sqrt(np.maximum(0.0, 1 - z * z))
            phi = 2 * np.pi * u[:, 1]
            d = np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=-1)
            pts = pts + 0.5 * self.jitter * self.spacing * d
        keep = self.inside(pts) if self.inside is not None else np.ones(len(pts), bool)
        pts = pts[keep]
        if self.max_particles is not None:
            pts = pts[: self.max_particles]
        builder.AddParticles(pts, np.broadcast_to(self.init_vel, pts.shape))
        return len(pts)


def box_inside_reference(center, size):
    """Shape::SignedDistance(p) <= 0 for an axis-aligned box exactly as the reference decides it (box.cpp:74-109 with
    Inside(point, bounds), geometry.h:1773-1783): inside, or within 1e-6 of ANY face plane."""
    c = np.asarray(center, float)
    h = np.asarray(size, float) / 2

    def f(p):
        q = p - c
        exact = np.all((q >= -h) & (q <= h), axis=-1)
        near = np.minimum(np.abs(-h - q), np.abs(h - q)) < 1e-6
        return exact | near.any(axis=-1)
    return f


def box_inside(center, size):
    c = np.asarray(center, float)
    h = np.asarray(size, float) / 2

    def f(p):
        return np.all(np.abs(p - c) <= h, axis=-1)
    return f
