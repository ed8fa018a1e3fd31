"""Thin Python view of the bbx C ABI, used by the tests and bench.py.

Names follow the reference's C++ API (PciSphSolver3::Setup/Advance, ColliderSetBuilder3, MakeBox,
UtilBuildGridForDomain ...; see include/bbx.h for the file:line of each).  The C++ facade with the
same class names for a C++ host lives in bubbles_b200/host/bubbles_api.h.  Nothing here computes:
every method is one call into libbbx.so.
"""
import ctypes as C

import numpy as np

from . import _lib as L


class BbxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"bbx error {code}: {msg}")
        self.code = code


def _check(rc):
    if rc != L.OK:
        raise BbxError(rc, L.load().bbx_last_error().decode())


def _vec3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def identity():
    return np.eye(4)


def Translate(x, y, z):
    """Transform Translate(x, y, z): returns (m, mInv) with the exact analytic inverse (transform.cpp:286-292)."""
    m, mi = np.eye(4), np.eye(4)
    m[:3, 3] = (x, y, z)
    mi[:3, 3] = (-x, -y, -z)
    return m, mi


def _xf(t):
    if t is None:
        return np.eye(4), np.eye(4)
    if isinstance(t, tuple):
        return np.asarray(t[0], dtype=np.float64), np.asarray(t[1], dtype=np.float64)
    m = np.asarray(t, dtype=np.float64)
    return m, np.linalg.inv(m)


def _shape(kind, to_world, reverse, **kw):
    c = L.Collider()
    c.type = kind
    c.reverse_orientation = int(bool(reverse))
    c.active = 1
    m, mi = _xf(to_world)
    c.object_to_world[:] = m.ravel()
    c.world_to_object[:] = mi.ravel()
    for k, v in kw.items():
        if isinstance(v, (tuple, list, np.ndarray)):
            getattr(c, k)[:] = v
        else:
            setattr(c, k, v)
    return c


def MakeBox(to_world, size, reverse_orientation=False):
    return _shape(L.COLLIDER_BOX, to_world, reverse_orientation, size=tuple(float(s) for s in size))


def MakeSphere(to_world, radius, reverse_orientation=False):
    return _shape(L.COLLIDER_SPHERE, to_world, reverse_orientation, radius=float(radius))


def sdf_grid_layout(bounds_min, bounds_max, dx=0.01, margin=0.1):
    """Node layout of Shape::InitSDFShape (shape.h:201-230): returns (resolution nodes, spacing, origin)."""
    lo = np.asarray(bounds_min, dtype=np.float64).copy()
    hi = np.asarray(bounds_max, dtype=np.float64).copy()
    scale = np.abs(hi - lo)
    lo -= margin * scale
    hi += margin * scale
    width, height, depth = np.abs(hi - lo)
    res = int(np.ceil(width / dx))
    dx = width / float(res)
    res_y = int(np.ceil(res * height / width))
    res_z = int(np.ceil(res * depth / width))
    nodes = (res + 1, res_y + 1, res_z + 1)  # VertexCentered: n+1 nodes per axis
    return nodes, dx, lo


def MakeSDFShape(bounds_min, bounds_max, sdf, dx=0.01, margin=0.1):
    """MakeSDFShape (shape.h:281-286): bakes `sdf(points[n,3]) -> distances[n]` on the vertex grid."""
    nodes, dx, origin = sdf_grid_layout(bounds_min, bounds_max, dx, margin)
    ix, iy, iz = np.meshgrid(np.arange(nodes[0]), np.arange(nodes[1]), np.arange(nodes[2]), indexing="ij")
    pts = np.stack([origin[0] + dx * ix, origin[1] + dx * iy, origin[2] + dx * iz], axis=-1)
    field = np.asarray(sdf(pts.reshape(-1, 3)), dtype=np.float64).reshape(nodes)
    field = np.ascontiguousarray(field.transpose(2, 1, 0))  # x fastest in memory (LinearIndex)
    c = _shape(L.COLLIDER_SDF, None, False, sdf_resolution=nodes, sdf_spacing=(dx, dx, dx),
               sdf_origin=tuple(origin))
    c._field = field
    c.sdf_field = field.ctypes.data
    return c


def MakeMesh(vertices, triangles, sdf, reverse_orientation=False):
    """MakeMesh (shape.h:275) + the SDF grid the collider set generates for it (GenerateShapeSDF, shape.cpp:479-511):
    vertices [n, 3], triangles [m, 3], sdf = dict(res=node counts, spacing=(dx, dy, dz), origin, field float64 x-fastest)."""
    c = _shape(L.COLLIDER_MESH, None, reverse_orientation, sdf_resolution=tuple(int(r) for r in sdf["res"]),
               sdf_spacing=tuple(float(x) for x in sdf["spacing"]), sdf_origin=tuple(float(x) for x in sdf["origin"]))
    c._field = np.ascontiguousarray(sdf["field"], dtype=np.float64)
    c._points = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
    c._indices = np.ascontiguousarray(triangles, dtype=np.int32).reshape(-1, 3)
    c.sdf_field = c._field.ctypes.data
    c.mesh_vertices, c.mesh_triangles = len(c._points), len(c._indices)
    c.mesh_points, c.mesh_indices = c._points.ctypes.data, c._indices.ctypes.data
    return c


def UtilBuildGridForDomain(domain_min, domain_max, spacing, spacing_scale):
    g = L.GridDesc()
    _check(L.load().bbx_grid_for_domain(_vec3(domain_min), _vec3(domain_max), spacing, spacing_scale, C.byref(g)))
    return g


def MakeGrid(resolution, p0, p1):
    g = L.GridDesc()
    _check(L.load().bbx_grid_build((C.c_int * 3)(*resolution), _vec3(p0), _vec3(p1), C.byref(g)))
    return g


class ColliderSetBuilder3:
    def __init__(self):
        self.colliders = []

    def AddCollider3(self, shape, friction=0.0):
        shape.friction = float(friction)
        self.colliders.append(shape)

    def GetColliderSet(self):
        return list(self.colliders)


class Engine:
    """One bbx_engine handle."""

    def __init__(self, grid, spacing, kernel_scale, max_particles, device=0, with_gravity=True,
                 reference_compat=True, slab=None, **overrides):
        self.lib = L.load()
        cfg = L.Config()
        _check(self.lib.bbx_config_default(C.byref(cfg), int(with_gravity)))
        cfg.device = device
        cfg.max_particles = int(max_particles)
        cfg.spacing = spacing
        cfg.kernel_scale = kernel_scale
        cfg.pcisph_reference_compat = int(reference_compat)
        cfg.grid = grid
        if slab is not None:  # owned global cell planes [z0, z1) of a multi-GPU slab decomposition
            cfg.slab_z_begin, cfg.slab_z_end = int(slab[0]), int(slab[1])
        for k, v in overrides.items():
            if k == "gravity":
                cfg.gravity[:] = v
            else:
                setattr(cfg, k, v)
        self.cfg = cfg
        self.grid = grid
        h = L._E()
        _check(self.lib.bbx_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self._keep = []

    def close(self):
        if self.h:
            self.lib.bbx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setup
    @property
    def mass(self):
        m = C.c_double()
        _check(self.lib.bbx_get_mass(self.h, C.byref(m)))
        return m.value

    def delta(self, dt):
        d = C.c_double()
        _check(self.lib.bbx_get_delta(self.h, dt, C.byref(d)))
        return d.value

    @staticmethod
    def _arr(a):
        a = np.asarray(a)
        if a.dtype == np.float32:
            return np.ascontiguousarray(a), L.F32
        return np.ascontiguousarray(a, dtype=np.float64), L.F64

    @staticmethod
    def _pair(pos, vel, rows=None):
        """contiguous (pos, vel, dtype code) of equal shape [n, 3] (and exactly `rows` rows when given): the C ABI
        reads n x 3 values from each pointer, a short array would be read past its end"""
        pos, dt = Engine._arr(pos)
        vel = np.ascontiguousarray(vel, dtype=pos.dtype)
        pos, vel = pos.reshape(-1, 3), vel.reshape(-1, 3)
        if pos.shape != vel.shape:
            raise ValueError(f"positions {pos.shape} and velocities {vel.shape} differ in shape")
        if rows is not None and len(pos) != rows:
            raise ValueError(f"{len(pos)} rows given, the engine holds {rows} particles")
        return pos, vel, dt

    def set_particles(self, pos, vel):
        pos, vel, dt = self._pair(pos, vel)
        _check(self.lib.bbx_set_particles(self.h, len(pos), pos.ctypes.data, vel.ctypes.data, dt))

    def set_particles_ids(self, pos, vel, ids=None):
        """Slab engines: keeps the particles of the owned planes; collective over the slab group."""
        pos, vel, dt = self._pair(pos, vel)
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int32)
            if len(ids) != len(pos):
                raise ValueError(f"{len(ids)} ids for {len(pos)} particles")
        _check(self.lib.bbx_set_particles_ids(self.h, len(pos), pos.ctypes.data, vel.ctypes.data,
                                              None if ids is None else ids.ctypes.data, dt))

    def comm_init_local(self, rank, nranks, group):
        _check(self.lib.bbx_comm_init_local(self.h, rank, nranks, group.encode()))

    def comm_init(self, rank, nranks, unique_id):
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        _check(self.lib.bbx_comm_init(self.h, rank, nranks, buf))

    def append_particles(self, pos, vel):
        pos, vel, dt = self._pair(pos, vel)
        _check(self.lib.bbx_append_particles(self.h, len(pos), pos.ctypes.data, vel.ctypes.data, dt))

    def append_particles_ids(self, pos, vel, ids):
        """append with explicit global ids (slab engines: collective, each rank keeps the particles of its planes)"""
        pos, vel, dt = self._pair(pos, vel)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        if len(ids) != len(pos):
            raise ValueError(f"{len(ids)} ids for {len(pos)} particles")
        _check(self.lib.bbx_append_particles_ids(self.h, len(pos), pos.ctypes.data, vel.ctypes.data, ids.ctypes.data, dt))

    def overwrite_state(self, pos, vel):
        pos, vel, dt = self._pair(pos, vel, rows=self.n)
        _check(self.lib.bbx_overwrite_state(self.h, pos.ctypes.data, vel.ctypes.data, dt))

    @property
    def p2p(self):
        """True when halo results are stored straight into the neighbours' ghost slots (peer memory)."""
        v = C.c_int()
        _check(self.lib.bbx_halo_mode(self.h, C.byref(v)))
        return bool(v.value)

    def overwrite_owned(self, pos, vel):
        """Overwrite the owned particles, rows in the order of the last download_owned (collective on slabs)."""
        pos, vel, dt = self._pair(pos, vel, rows=self.n)
        _check(self.lib.bbx_overwrite_owned(self.h, pos.ctypes.data, vel.ctypes.data, dt))

    @property
    def n(self):
        n = C.c_int()
        _check(self.lib.bbx_particle_count(self.h, C.byref(n)))
        return n.value

    def set_colliders(self, colliders):
        arr = (L.Collider * max(1, len(colliders)))(*colliders)
        self._keep = [colliders, arr]
        _check(self.lib.bbx_set_colliders(self.h, len(colliders), arr))

    def set_collider_active(self, index, active):
        _check(self.lib.bbx_set_collider_active(self.h, index, int(active)))

    def collider_distance(self, index, points):
        """Shape::ClosestDistance of collider `index` at points [n, 3] (device evaluation)"""
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(len(pts), dtype=np.float64)
        _check(self.lib.bbx_collider_distance(self.h, int(index), len(pts), pts.ctypes.data, out.ctypes.data))
        return out

    def update_collider(self, index, collider):
        _check(self.lib.bbx_update_collider(self.h, index, C.byref(collider)))

    def set_param(self, param, value):
        """a solver constant the reference reads live (viscosity, pseudo-viscosity, compat flag ...): next sub-step on"""
        _check(self.lib.bbx_set_param(self.h, int(param), float(value)))

    # -- stepping
    def step_pcisph(self, dt):
        _check(self.lib.bbx_step_pcisph(self.h, dt))

    def step_sph(self, dt):
        _check(self.lib.bbx_step_sph(self.h, dt))

    def step_many(self, dt, n, solver=L.SOLVER_PCISPH):
        _check(self.lib.bbx_step_many(self.h, dt, solver, n))

    def step_many_timed(self, dt, n, solver=L.SOLVER_PCISPH):
        """n sub-steps; returns the device time (ms) of the whole region, events on the engine's stream."""
        ms = C.c_float()
        _check(self.lib.bbx_step_many_timed(self.h, dt, solver, n, C.byref(ms)))
        return ms.value

    def advance(self, seconds, solver=L.SOLVER_PCISPH):
        sub, ms = C.c_int(), C.c_float()
        _check(self.lib.bbx_advance(self.h, seconds, solver, C.byref(sub), C.byref(ms)))
        return sub.value, ms.value

    def run_phase(self, phase, dt):
        _check(self.lib.bbx_run_phase(self.h, phase, dt))

    def synchronize(self):
        _check(self.lib.bbx_synchronize(self.h))

    def set_timing(self, on):
        _check(self.lib.bbx_set_timing(self.h, int(on)))

    def stats(self):
        s = L.StepStats()
        _check(self.lib.bbx_stats(self.h, C.byref(s)))
        return s

    # -- results
    def download(self, field, dtype=np.float64):
        n = self.n
        if field == L.NEIGHBOR_COUNT:
            out = np.zeros(n, dtype=np.int32)
            code = L.I32
        else:
            vec = field in (L.POSITION, L.VELOCITY, L.FORCE, L.PRED_POSITION, L.PRESSURE_FORCE, L.FORCE_NP)
            out = np.zeros((n, 3) if vec else n, dtype=dtype)
            code = L.F64 if out.dtype == np.float64 else L.F32
        _check(self.lib.bbx_download(self.h, field, out.ctypes.data, code))
        return out

    _VEC = (L.POSITION, L.VELOCITY, L.FORCE, L.PRED_POSITION, L.PRESSURE_FORCE, L.FORCE_NP)

    def download_owned(self, field=None, dtype=np.float64):
        """(ids, values) of the owned particles in the engine's cell order; field None: ids only."""
        n = self.n
        ids = np.zeros(n, dtype=np.int32)
        cnt = C.c_int()
        if field is None:
            _check(self.lib.bbx_download_owned(self.h, 0, None, L.F32, ids.ctypes.data, C.byref(cnt)))
            return ids, None
        if field == L.NEIGHBOR_COUNT:
            out = np.zeros(n, dtype=np.int32)
            code = L.I32
        else:
            out = np.zeros((n, 3) if field in self._VEC else n, dtype=dtype)
            code = L.F64 if out.dtype == np.float64 else L.F32
        _check(self.lib.bbx_download_owned(self.h, field, out.ctypes.data, code, ids.ctypes.data, C.byref(cnt)))
        return ids, out

    def export_neighbors_owned(self):
        n = self.n
        counts = np.zeros(n, dtype=np.int32)
        ids = np.full((n, L.MAX_NEIGHBORS), -1, dtype=np.int32)
        _check(self.lib.bbx_export_neighbors_owned(self.h, counts.ctypes.data, ids.ctypes.data))
        return counts, ids

    def plane_counts(self):
        """owned particles per GLOBAL cell plane (bbx_plane_counts)"""
        out = (C.c_longlong * self.grid.n[2])()
        _check(self.lib.bbx_plane_counts(self.h, out))
        return np.array(out[:], dtype=np.int64)

    def rebalance(self, z_bounds):
        """collective over the slab group: move to the cuts z_bounds (bbx_rebalance)"""
        zb = (C.c_int * len(z_bounds))(*[int(z) for z in z_bounds])
        _check(self.lib.bbx_rebalance(self.h, zb))

    def export_cells(self):
        count = np.zeros(self.grid.total, dtype=np.int32)
        order = np.zeros(self.n, dtype=np.int32)
        _check(self.lib.bbx_export_cells(self.h, count.ctypes.data, order.ctypes.data))
        return count, order

    def export_neighbors(self):
        n = self.n
        counts = np.zeros(n, dtype=np.int32)
        ids = np.full((n, L.MAX_NEIGHBORS), -1, dtype=np.int32)
        _check(self.lib.bbx_export_neighbors(self.h, counts.ctypes.data, ids.ctypes.data))
        return counts, ids

    def query_cells(self, cells, points, d):
        """MapGridEmit's per-cell test on the device: (chain length of each point's cell, blocked flags)"""
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        if len(cells) != len(points):
            raise ValueError("one cell id per point")
        size = np.zeros(len(cells), dtype=np.int32)
        blocked = np.zeros(len(cells), dtype=np.int32)
        _check(self.lib.bbx_query_cells(self.h, len(cells), cells.ctypes.data, points.ctypes.data, float(d), size.ctypes.data, blocked.ctypes.data))
        return size, blocked

    def inject_chains(self, cell_count, cell_order):
        cc = np.ascontiguousarray(cell_count, dtype=np.int32)
        co = np.ascontiguousarray(cell_order, dtype=np.int32)
        _check(self.lib.bbx_inject_chains(self.h, cc.ctypes.data, co.ctypes.data))

    def set_rebuild_flag(self, flag):
        _check(self.lib.bbx_set_rebuild_flag(self.h, int(flag)))

    @property
    def launches(self):
        c = C.c_longlong()
        _check(self.lib.bbx_launch_count(self.h, C.byref(c)))
        return c.value

    def kernel_time(self, phase):
        ms, k = C.c_float(), C.c_int()
        _check(self.lib.bbx_kernel_time(self.h, phase, C.byref(ms), C.byref(k)))
        return ms.value, k.value

    def reset_kernel_time(self):
        _check(self.lib.bbx_reset_kernel_time(self.h))
