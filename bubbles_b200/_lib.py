"""ctypes binding of the bbx C ABI (include/bbx.h).  Plumbing only: every call goes to
bubbles_b200/lib/libbbx.so (hand-written CUDA for sm_100a).  There is no Python or CPU
implementation behind these functions -- if the library is missing, importing this module raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# BBX_LIB: development aid (A/B runs of kernel variants built from the same sources); default = the in-tree build
LIB_PATH = os.environ.get("BBX_LIB") or os.path.join(HERE, "lib", "libbbx.so")

MAX_NEIGHBORS = 100
MAX_COLLIDERS = 16

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_CAPACITY, ERR_OUT_OF_DOMAIN, ERR_COMM = range(7)
SOLVER_PCISPH, SOLVER_SPH = 0, 1
(POSITION, VELOCITY, FORCE, DENSITY, PRESSURE, PRED_POSITION, PRED_DENSITY, PRESSURE_FORCE, FORCE_NP,
 DENSITY_ERROR, NEIGHBOR_COUNT) = range(11)
F32, F64, I32 = 0, 1, 2
(PHASE_GRID, PHASE_DENSITY, PHASE_FORCE_NP, PHASE_PREDICT, PHASE_PRESSURE, PHASE_PRESSURE_FORCE,
 PHASE_INTEGRATE) = range(7)
COLLIDER_BOX, COLLIDER_SPHERE, COLLIDER_SDF, COLLIDER_MESH = 0, 1, 2, 3
(PARAM_VISCOSITY, PARAM_PSEUDO_VISCOSITY, PARAM_DRAG, PARAM_RESTITUTION, PARAM_NEGATIVE_PRESSURE_SCALE, PARAM_REFERENCE_COMPAT,
 PARAM_MAX_ITERATIONS, PARAM_MAX_DENSITY_ERROR_RATIO, PARAM_TIME_STEP_LIMIT_SCALE, PARAM_GRAVITY_X, PARAM_GRAVITY_Y, PARAM_GRAVITY_Z) = range(12)


class GridDesc(C.Structure):
    _fields_ = [("min", C.c_double * 3), ("max", C.c_double * 3), ("cell_len", C.c_double * 3),
                ("n", C.c_int * 3), ("total", C.c_int)]


class Collider(C.Structure):
    _fields_ = [("type", C.c_int), ("reverse_orientation", C.c_int), ("active", C.c_int), ("reserved", C.c_int),
                ("object_to_world", C.c_double * 16), ("world_to_object", C.c_double * 16),
                ("size", C.c_double * 3), ("radius", C.c_double), ("friction", C.c_double),
                ("linear_velocity", C.c_double * 3), ("angular_velocity", C.c_double * 3),
                ("sdf_resolution", C.c_int * 3), ("reserved2", C.c_int),
                ("sdf_spacing", C.c_double * 3), ("sdf_origin", C.c_double * 3),
                ("sdf_field", C.c_void_p),
                ("mesh_vertices", C.c_int), ("mesh_triangles", C.c_int), ("mesh_points", C.c_void_p), ("mesh_indices", C.c_void_p)]


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_int), ("device", C.c_int), ("max_particles", C.c_int),
                ("pcisph_max_iterations", C.c_int), ("pcisph_reference_compat", C.c_int),
                ("with_gravity", C.c_int),
                ("spacing", C.c_double), ("kernel_scale", C.c_double), ("target_density", C.c_double),
                ("viscosity", C.c_double), ("drag", C.c_double), ("eos_exponent", C.c_double),
                ("sound_speed", C.c_double), ("negative_pressure_scale", C.c_double),
                ("pseudo_viscosity", C.c_double), ("gravity", C.c_double * 3),
                ("pcisph_max_density_error_ratio", C.c_double), ("restitution", C.c_double),
                ("time_step_limit_scale", C.c_double), ("grid", GridDesc),
                ("slab_z_begin", C.c_int), ("slab_z_end", C.c_int),
                ("ghost_capacity", C.c_int), ("reserved", C.c_int)]


class StepStats(C.Structure):
    _fields_ = [("particles", C.c_int), ("ghosts", C.c_int), ("substeps", C.c_int),
                ("pcisph_iterations", C.c_int), ("full_rebuild", C.c_int), ("rebuild_flag", C.c_int),
                ("neighbor_overflow", C.c_int), ("lost_particles", C.c_int), ("clamped", C.c_int),
                ("nan_count", C.c_int), ("max_force", C.c_float), ("max_density_error", C.c_float),
                ("ms_grid", C.c_float), ("ms_step", C.c_float),
                ("exact_passes", C.c_int), ("max_candidates", C.c_int), ("occupied_cells", C.c_int), ("unstaged_tiles", C.c_int)]


# every symbol include/bbx.h declares: name -> (restype, argtypes)
_E = C.c_void_p
SYMBOLS = {
    "bbx_last_error": (C.c_char_p, []),
    "bbx_version": (C.c_int, []),
    "bbx_config_default": (C.c_int, [C.POINTER(Config), C.c_int]),
    "bbx_grid_for_domain": (C.c_int, [C.c_double * 3, C.c_double * 3, C.c_double, C.c_double, C.POINTER(GridDesc)]),
    "bbx_grid_build": (C.c_int, [C.c_int * 3, C.c_double * 3, C.c_double * 3, C.POINTER(GridDesc)]),
    "bbx_create": (C.c_int, [C.POINTER(Config), C.POINTER(_E)]),
    "bbx_destroy": (C.c_int, [_E]),
    "bbx_set_param": (C.c_int, [_E, C.c_int, C.c_double]),
    "bbx_get_mass": (C.c_int, [_E, C.POINTER(C.c_double)]),
    "bbx_get_delta": (C.c_int, [_E, C.c_double, C.POINTER(C.c_double)]),
    "bbx_set_particles": (C.c_int, [_E, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "bbx_set_particles_ids": (C.c_int, [_E, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "bbx_append_particles": (C.c_int, [_E, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "bbx_append_particles_ids": (C.c_int, [_E, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "bbx_particle_count": (C.c_int, [_E, C.POINTER(C.c_int)]),
    "bbx_overwrite_state": (C.c_int, [_E, C.c_void_p, C.c_void_p, C.c_int]),
    "bbx_overwrite_owned": (C.c_int, [_E, C.c_void_p, C.c_void_p, C.c_int]),
    "bbx_set_colliders": (C.c_int, [_E, C.c_int, C.POINTER(Collider)]),
    "bbx_update_collider": (C.c_int, [_E, C.c_int, C.POINTER(Collider)]),
    "bbx_set_collider_active": (C.c_int, [_E, C.c_int, C.c_int]),
    "bbx_collider_distance": (C.c_int, [_E, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "bbx_step_pcisph": (C.c_int, [_E, C.c_double]),
    "bbx_step_sph": (C.c_int, [_E, C.c_double]),
    "bbx_advance": (C.c_int, [_E, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float)]),
    "bbx_step_many": (C.c_int, [_E, C.c_double, C.c_int, C.c_int]),
    "bbx_step_many_timed": (C.c_int, [_E, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "bbx_run_phase": (C.c_int, [_E, C.c_int, C.c_double]),
    "bbx_synchronize": (C.c_int, [_E]),
    "bbx_set_timing": (C.c_int, [_E, C.c_int]),
    "bbx_stats": (C.c_int, [_E, C.POINTER(StepStats)]),
    "bbx_download": (C.c_int, [_E, C.c_int, C.c_void_p, C.c_int]),
    "bbx_download_owned": (C.c_int, [_E, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]),
    "bbx_download_state": (C.c_int, [_E, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "bbx_export_cells": (C.c_int, [_E, C.c_void_p, C.c_void_p]),
    "bbx_export_neighbors": (C.c_int, [_E, C.c_void_p, C.c_void_p]),
    "bbx_export_neighbors_owned": (C.c_int, [_E, C.c_void_p, C.c_void_p]),
    "bbx_query_cells": (C.c_int, [_E, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "bbx_inject_chains": (C.c_int, [_E, C.c_void_p, C.c_void_p]),
    "bbx_set_rebuild_flag": (C.c_int, [_E, C.c_int]),
    "bbx_launch_count": (C.c_int, [_E, C.POINTER(C.c_longlong)]),
    "bbx_kernel_time": (C.c_int, [_E, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "bbx_reset_kernel_time": (C.c_int, [_E]),
    "bbx_comm_unique_id": (C.c_int, [C.c_void_p]),
    "bbx_comm_init": (C.c_int, [_E, C.c_int, C.c_int, C.c_void_p]),
    "bbx_comm_init_local": (C.c_int, [_E, C.c_int, C.c_int, C.c_char_p]),
    "bbx_halo_mode": (C.c_int, [_E, C.POINTER(C.c_int)]),
    "bbx_slab_plan": (C.c_int, [C.c_int, C.POINTER(C.c_longlong), C.c_int, C.POINTER(C.c_int)]),
    "bbx_plane_counts": (C.c_int, [_E, C.POINTER(C.c_longlong)]),
    "bbx_rebalance": (C.c_int, [_E, C.POINTER(C.c_int)]),
    "bbx_slab_plan_step": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "bbx_plane_histogram": (C.c_int, [C.POINTER(GridDesc), C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_longlong)]),
}

_lib = None


def load():
    """Load libbbx.so and bind every declared symbol.  Fails loudly when the CUDA library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). bubbles_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
