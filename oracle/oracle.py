"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/bbx_oracle.c (the CPU restatement of the
reference's 3D PCISPH/SPH sub-step) plus a driver for oracle/_ref/bbref (the unmodified reference,
built by oracle/build_ref.sh).  Imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package never imports it.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "bbref")
MAX_BUCKET = 100


def build(force=False):
    """Compile the C restatement (gcc, no fast-math, no FMA contraction)."""
    src = os.path.join(HERE, "bbx_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-w",
                               "-o", LIB_PATH, src, "-lm"])
    return LIB_PATH


class Grid(C.Structure):
    _fields_ = [("gmin", C.c_double * 3), ("gmax", C.c_double * 3), ("glen", C.c_double * 3),
                ("gn", C.c_int * 3), ("total", C.c_int)]


class Collider(C.Structure):
    _fields_ = [("type", C.c_int), ("reverse", C.c_int), ("active", C.c_int), ("_pad", C.c_int),
                ("o2w", C.c_double * 16), ("o2w_inv", C.c_double * 16),
                ("w2o", C.c_double * 16), ("w2o_inv", C.c_double * 16),
                ("size", C.c_double * 3), ("radius", C.c_double), ("friction", C.c_double),
                ("linvel", C.c_double * 3), ("angvel", C.c_double * 3),
                ("sdf_res", C.c_int * 3), ("_pad2", C.c_int),
                ("sdf_spacing", C.c_double * 3), ("sdf_origin", C.c_double * 3),
                ("sdf_field", C.c_void_p),
                ("n_vertices", C.c_int), ("n_triangles", C.c_int), ("vertices", C.c_void_p), ("indices", C.c_void_p),
                ("mesh_min", C.c_double * 3), ("mesh_max", C.c_double * 3)]


class Params(C.Structure):
    _fields_ = [("spacing", C.c_double), ("h", C.c_double), ("rho0", C.c_double), ("mass", C.c_double),
                ("viscosity", C.c_double), ("drag", C.c_double), ("eos_exponent", C.c_double),
                ("sound_speed", C.c_double), ("neg_pressure_scale", C.c_double),
                ("pseudo_viscosity", C.c_double), ("gravity", C.c_double * 3),
                ("delta_denom", C.c_double), ("mass_over_rho0_sq", C.c_double),
                ("grid", Grid), ("n_colliders", C.c_int), ("_pad", C.c_int),
                ("colliders", C.POINTER(Collider))]


class State(C.Structure):
    _fields_ = [("n", C.c_int), ("rebuild_flag", C.c_int)] + \
        [(k, C.c_void_p) for k in ("pos", "vel", "force", "density", "pressure", "pos_pred", "vel_pred",
                                   "force_p", "density_pred", "cell_count", "cell_order", "cell_count2",
                                   "cell_order2", "nbr_count", "nbr_ids")] + \
        [("lost", C.c_int), ("overflow", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        assert L.orc_sizeof_params() == C.sizeof(Params), (L.orc_sizeof_params(), C.sizeof(Params))
        assert L.orc_sizeof_collider() == C.sizeof(Collider)
        assert L.orc_sizeof_state() == C.sizeof(State)
        for name in ("orc_compute_mass", "orc_delta_denom", "orc_delta", "orc_w_std", "orc_w_spiky",
                     "orc_dw_spiky", "orc_d2w_spiky"):
            getattr(L, name).restype = C.c_double
        L.orc_compute_mass.argtypes = [C.c_double] * 3
        L.orc_delta_denom.argtypes = [C.c_double] * 2
        L.orc_delta.argtypes = [C.POINTER(Params), C.c_double]
        for name in ("orc_w_std", "orc_w_spiky", "orc_dw_spiky", "orc_d2w_spiky"):
            getattr(L, name).argtypes = [C.c_double, C.c_double]
        L.orc_number_of_time_steps.restype = C.c_uint
        L.orc_number_of_time_steps.argtypes = [C.POINTER(Params), C.c_int, C.c_void_p, C.c_double, C.c_double]
        L.orc_pcisph_substep.argtypes = [C.POINTER(Params), C.POINTER(State), C.c_double, C.c_int, C.c_int,
                                         C.c_double]
        L.orc_sph_substep.argtypes = [C.POINTER(Params), C.POINTER(State), C.c_double]
        L.orc_sph_substep_gs.argtypes = [C.POINTER(Params), C.POINTER(State), C.c_double]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


IDENTITY = np.eye(4)


def translate(x, y, z):
    m = np.eye(4)
    m[:3, 3] = (x, y, z)
    return m


def make_collider(kind, o2w=IDENTITY, size=(0, 0, 0), radius=0.0, reverse=False, friction=0.0,
                  active=True, linvel=(0, 0, 0), angvel=(0, 0, 0), sdf=None, w2o=None, mesh=None):
    """kind: 'box' | 'sphere' | 'sdf' | 'mesh'. sdf = dict(res=(nx,ny,nz) nodes, spacing, origin, field float64);
    mesh = (vertices [n, 3] float64, triangles [m, 3] int32), together with the sdf generated for it."""
    c = Collider()
    c.type = {"box": 0, "sphere": 1, "sdf": 2, "mesh": 3}[kind]
    c.reverse = int(reverse)
    c.active = int(active)
    o2w = np.asarray(o2w, dtype=np.float64)
    if w2o is None:
        w2o = np.linalg.inv(o2w)
        # Translate()/Scale() in the reference carry an exact analytic inverse (transform.cpp:286-306)
        if np.allclose(o2w[:3, :3], np.eye(3), atol=0, rtol=0):
            w2o = np.eye(4)
            w2o[:3, 3] = -o2w[:3, 3]
    c.o2w[:] = o2w.ravel()
    c.o2w_inv[:] = np.asarray(w2o).ravel()
    c.w2o[:] = np.asarray(w2o).ravel()
    c.w2o_inv[:] = o2w.ravel()
    c.size[:] = size
    c.radius = radius
    c.friction = friction
    c.linvel[:] = linvel
    c.angvel[:] = angvel
    if sdf is not None:
        c.sdf_res[:] = sdf["res"]
        c.sdf_spacing[:] = sdf["spacing"]
        c.sdf_origin[:] = sdf["origin"]
        c._field = np.ascontiguousarray(sdf["field"], dtype=np.float64)
        c.sdf_field = c._field.ctypes.data
    if mesh is not None:
        c._vertices = np.ascontiguousarray(mesh[0], dtype=np.float64)
        c._indices = np.ascontiguousarray(mesh[1], dtype=np.int32)
        c.n_vertices, c.n_triangles = len(c._vertices), len(c._indices)
        c.vertices, c.indices = c._vertices.ctypes.data, c._indices.ctypes.data
        used = c._vertices[np.unique(c._indices)]      # bounds of the BVH root = of the triangles (bvh.cpp, MeshGetBounds)
        c.mesh_min[:] = used.min(axis=0)
        c.mesh_max[:] = used.max(axis=0)
    return c


class Oracle:
    """Holds orc_params + orc_state for one scene."""

    def __init__(self, spacing, scale, domain_min, domain_max, colliders, rho0=1000.0, gravity=True,
                 viscosity=None, drag=None):
        L = lib()
        self.L = L
        self.P = Params()
        L.orc_default_params(C.byref(self.P), int(gravity))
        L.orc_setup_scalars(C.byref(self.P), C.c_double(spacing), C.c_double(scale), C.c_double(rho0))
        dmin = (C.c_double * 3)(*domain_min)
        dmax = (C.c_double * 3)(*domain_max)
        L.orc_grid_for_domain(dmin, dmax, C.c_double(spacing), C.c_double(scale), C.byref(self.P.grid))
        if viscosity is not None:
            self.P.viscosity = viscosity
        if drag is not None:
            self.P.drag = drag
        self._colliders = (Collider * max(1, len(colliders)))(*colliders)
        self._collider_objs = colliders
        self.P.colliders = self._colliders
        self.P.n_colliders = len(colliders)
        self.S = None

    # grid facts
    @property
    def grid(self):
        g = self.P.grid
        return dict(min=np.array(g.gmin[:]), max=np.array(g.gmax[:]), len=np.array(g.glen[:]),
                    n=np.array(g.gn[:]), total=g.total)

    def hash(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        self.L.orc_hash.argtypes = [C.POINTER(Grid), C.c_void_p]
        return np.array([self.L.orc_hash(C.byref(self.P.grid), _p(pos[i])) for i in range(len(pos))],
                        dtype=np.int32)

    def set_particles(self, pos, vel):
        n = len(pos)
        a = {}
        a["pos"] = np.ascontiguousarray(pos, dtype=np.float64).copy()
        a["vel"] = np.ascontiguousarray(vel, dtype=np.float64).copy()
        for k in ("force", "pos_pred", "vel_pred", "force_p"):
            a[k] = np.zeros((n, 3))
        for k in ("density", "pressure", "density_pred"):
            a[k] = np.zeros(n)
        total = self.P.grid.total
        for k in ("cell_count", "cell_count2"):
            a[k] = np.zeros(total, dtype=np.int32)
        for k in ("cell_order", "cell_order2", "nbr_count"):
            a[k] = np.zeros(n, dtype=np.int32)
        a["nbr_ids"] = np.full((n, MAX_BUCKET), -1, dtype=np.int32)
        self.a = a
        S = State()
        S.n = n
        S.rebuild_flag = 0
        for k, v in a.items():
            setattr(S, k, v.ctypes.data)
        self.S = S
        # PciSphSolver3::Setup: initial serial DistributeByParticle (pcisph_solver3.cpp:140-143)
        self.L.orc_full_rebuild(C.byref(self.P.grid), n, _p(a["pos"]), _p(a["cell_count"]), _p(a["cell_order"]))

    def append_particles(self, pos, vel):
        """ContinuousParticleSetBuilder3::AddParticle + Commit (src/core/grid.h:1409-1441): ids continue from n, the
        new particles join the tail of their cells' chains; everything else of the old state is kept."""
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        vel = np.ascontiguousarray(vel, dtype=np.float64).reshape(-1, 3)
        k, n = len(pos), self.S.n
        if k == 0:
            return
        old = {name: self.arr(name) for name in self.a}   # current roles (the chain buffers may be swapped)
        a = {}
        for name in ("pos", "vel", "force", "pos_pred", "vel_pred", "force_p"):
            a[name] = np.zeros((n + k, 3)); a[name][:n] = old[name]
        a["pos"][n:] = pos; a["vel"][n:] = vel
        for name in ("density", "pressure", "density_pred"):
            a[name] = np.zeros(n + k); a[name][:n] = old[name]
        total = self.P.grid.total
        a["cell_count"] = np.zeros(total, dtype=np.int32); a["cell_count2"] = np.zeros(total, dtype=np.int32)
        a["cell_order"] = np.zeros(n + k, dtype=np.int32); a["cell_order2"] = np.zeros(n + k, dtype=np.int32)
        self.L.orc_append_chains(C.byref(self.P.grid), n, k, _p(pos), _p(old["cell_count"]), _p(old["cell_order"]),
                                 _p(a["cell_count"]), _p(a["cell_order"]))
        a["nbr_count"] = np.zeros(n + k, dtype=np.int32); a["nbr_count"][:n] = old["nbr_count"]
        a["nbr_ids"] = np.full((n + k, MAX_BUCKET), -1, dtype=np.int32); a["nbr_ids"][:n] = old["nbr_ids"]
        self.a = a
        self.S.n = n + k
        for name, v in a.items():
            setattr(self.S, name, v.ctypes.data)

    def map_grid(self):
        """ContinuousParticleSetBuilder3::MapGrid (src/core/grid.h:1288-1321): remember, for every occupied cell, the
        positions of its chain (in chain order) -- the emission template of map_grid_emit."""
        cc, co = self.arr("cell_count"), self.arr("cell_order")
        start = np.concatenate([[0], np.cumsum(cc)])
        self.mapped = {int(c): self.a["pos"][co[start[c]:start[c + 1]]].copy() for c in np.nonzero(cc)[0]}

    def map_grid_emit(self, velocity, d=0.02):
        """ContinuousParticleSetBuilder3::MapGridEmit (src/core/grid.h:1367-1407): for every mapped cell (ascending id)
        whose CURRENT chain has room (< 100), re-emit its first min(100 - size, len) template positions that have no
        particle of the current chain within d; one Commit at the end.  Returns the number of particles added."""
        cc, co = self.arr("cell_count"), self.arr("cell_order")
        start = np.concatenate([[0], np.cumsum(cc)])
        pos = self.a["pos"]
        new = []
        for c in sorted(self.mapped):
            size = int(cc[c])
            if size >= MAX_BUCKET:
                continue
            tmpl = self.mapped[c]
            members = pos[co[start[c]:start[c + 1]]]
            for i in range(min(MAX_BUCKET - size, len(tmpl))):
                pi = tmpl[i]
                ok = True
                for pj in members:
                    dx = pj - pi          # Distance(pj, pi) = sqrt(|pj - pi|^2) (geometry.h:632-633, 797-800)
                    if np.sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]) < d:
                        ok = False
                        break
                if ok:
                    new.append(pi)
        if new:
            new = np.array(new)
            self.append_particles(new, np.tile(np.asarray(velocity, dtype=np.float64), (len(new), 1)))
        return len(new)

    def arr(self, name):
        """numpy view of a state array (follows the chain double-buffer swap)."""
        ptr = getattr(self.S, name)
        for v in self.a.values():
            if v.ctypes.data == ptr:
                return v
        raise KeyError(name)

    def set_chains(self, cell_count, cell_order):
        self.arr("cell_count")[:] = cell_count
        self.arr("cell_order")[:] = cell_order
        self.S.rebuild_flag = 0

    def delta(self, dt):
        return self.L.orc_delta(C.byref(self.P), dt)

    def substep_pcisph(self, dt, compat=True, max_it=5, max_err_ratio=0.01):
        return self.L.orc_pcisph_substep(C.byref(self.P), C.byref(self.S), dt, int(compat), max_it, max_err_ratio)

    def substep_sph(self, dt):
        """AdvanceTimeStep(SphSolver3*) in the race-free Jacobi form the engine implements"""
        self.L.orc_sph_substep(C.byref(self.P), C.byref(self.S), dt)

    def substep_sph_gs(self, dt):
        """the same sub-step exactly as the reference's CPU path runs it on ONE thread: particles in ascending id, each
        integrated inside its force evaluation (Gauss-Seidel) -- pinned bit-exactly against bbref (sph_run.npz)"""
        self.L.orc_sph_substep_gs(C.byref(self.P), C.byref(self.S), dt)

    def trace_pcisph(self, dt):
        """One compat sub-step phase by phase; returns dict of per-phase arrays (copies)."""
        L, P, S, a = self.L, self.P, self.S, self.a
        n = S.n
        out = {"rebuild_flag": int(S.rebuild_flag)}
        L.orc_update_grid(C.byref(P), C.byref(S))
        S.rebuild_flag = 0
        out["cell_count"] = self.arr("cell_count").copy()
        out["cell_order"] = self.arr("cell_order").copy()
        out["nbr_count"] = a["nbr_count"].copy()
        out["nbr_ids"] = a["nbr_ids"].copy()
        out["lost"], out["overflow"] = S.lost, S.overflow
        L.orc_density(C.byref(P), n, _p(a["pos"]), _p(a["nbr_count"]), _p(a["nbr_ids"]), _p(a["density"]),
                      _p(a["pressure"]))
        out["density"] = a["density"].copy()
        out["eos_pressure"] = a["pressure"].copy()
        L.orc_force_np(C.byref(P), n, _p(a["pos"]), _p(a["vel"]), _p(a["density"]), _p(a["nbr_count"]),
                       _p(a["nbr_ids"]), _p(a["force"]))
        out["force_np"] = a["force"].copy()
        delta = self.delta(dt)
        out["delta"] = delta
        a["density_pred"][:] = a["density"]
        a["force_p"][:] = 0
        a["pressure"][:] = 0
        L.orc_predict(C.byref(P), n, C.c_double(dt), _p(a["pos"]), _p(a["vel"]), _p(a["force"]), _p(a["force_p"]),
                      _p(a["pos_pred"]), _p(a["vel_pred"]))
        out["pos_pred"] = a["pos_pred"].copy()
        err = np.zeros(n)
        L.orc_pred_pressure(C.byref(P), n, C.c_double(delta), _p(a["pos_pred"]), _p(a["nbr_count"]),
                            _p(a["nbr_ids"]), _p(a["pressure"]), _p(a["density_pred"]), _p(err))
        out["density_pred"] = a["density_pred"].copy()
        out["pressure"] = a["pressure"].copy()
        out["density_error"] = err
        L.orc_pressure_force(C.byref(P), n, _p(a["pos"]), _p(a["pressure"]), _p(a["density_pred"]),
                             _p(a["nbr_count"]), _p(a["nbr_ids"]), _p(a["force_p"]))
        out["force_p"] = a["force_p"].copy()
        S.rebuild_flag = L.orc_integrate(C.byref(P), n, C.c_double(dt), _p(a["pos"]), _p(a["vel"]), _p(a["force"]),
                                         _p(a["force_p"]))
        L.orc_pseudo_viscosity(C.byref(P), n, C.c_double(dt), _p(a["pos"]), _p(a["vel"]), _p(a["density"]),
                               _p(a["nbr_count"]), _p(a["nbr_ids"]))
        out["pos_out"] = a["pos"].copy()
        out["vel_out"] = a["vel"].copy()
        out["force_out"] = a["force"].copy()
        out["rebuild_flag_out"] = int(S.rebuild_flag)
        return out

    def resolve_collision(self, pos, vel, radius, restitution):
        pos = np.ascontiguousarray(pos, dtype=np.float64).copy()
        vel = np.ascontiguousarray(vel, dtype=np.float64).copy()
        hit = np.zeros(len(pos), dtype=np.int32)
        self.L.orc_resolve_collision_many(C.byref(self.P), C.c_double(radius), C.c_double(restitution),
                                          len(pos), _p(pos), _p(vel), _p(hit))
        return pos, vel, hit

    def mesh_closest_distance(self, index, points):
        """Shape::MeshClosestDistance of collider `index` at points [n, 3]"""
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        self.L.orc_mesh_closest_distance.restype = C.c_double
        self.L.orc_mesh_closest_distance.argtypes = [C.POINTER(Collider), C.c_void_p]
        return np.array([self.L.orc_mesh_closest_distance(C.byref(self._colliders[index]), _p(pts[i])) for i in range(len(pts))])

    def number_of_time_steps(self, time_step, scale):
        return self.L.orc_number_of_time_steps(C.byref(self.P), self.S.n, _p(self.a["force"]), time_step, scale)


def bcc_points(bmin, bmax, spacing):
    """BccLatticePointGenerator (src/generator/bcclattice.cpp:5-36) through the C restatement."""
    L = lib()
    lo = (C.c_double * 3)(*bmin)
    hi = (C.c_double * 3)(*bmax)
    L.orc_bcc_points.argtypes = [C.c_double * 3, C.c_double * 3, C.c_double, C.c_void_p, C.c_int]
    n = L.orc_bcc_points(lo, hi, spacing, None, 1 << 30)
    pts = np.zeros((n, 3))
    L.orc_bcc_points(lo, hi, spacing, _p(pts), n)
    return pts


# ------------------------------------------------------------------ reference driver (oracle/_ref/bbref)

def ref_available():
    return os.path.exists(REF_BIN)


REF_GPU_BIN = os.path.join(os.path.dirname(REF_BIN), "bbref_gpu")


def ref_gpu_available():
    """The same driver linked for the reference's GPU path (its native mode); needs a GPU at run time."""
    return os.path.exists(REF_GPU_BIN)


def mat_str(m):
    return " ".join(repr(float(x)) for x in np.asarray(m, dtype=np.float64).ravel())


def write_particles(path, pos, vel):
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    vel = np.ascontiguousarray(vel, dtype=np.float64)
    with open(path, "wb") as f:
        f.write(np.int64(len(pos)).tobytes())
        f.write(pos.tobytes())
        f.write(vel.tobytes())


def run_ref(job_lines, workdir=None, timeout=3600, gpu=False):
    """Run bbref (gpu=True: bbref_gpu, the reference's GPU path) on a job; returns (stdout, workdir). Job lines may
    use {wd} for the work directory."""
    wd = workdir or tempfile.mkdtemp(prefix="bbref_")
    job = os.path.join(wd, "job.txt")
    with open(job, "w") as f:
        f.write("\n".join(l.format(wd=wd) for l in job_lines) + "\n")
    out = subprocess.run([REF_GPU_BIN if gpu else REF_BIN, job], check=True, capture_output=True, text=True, timeout=timeout).stdout
    return out, wd
