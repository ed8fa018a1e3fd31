/* bbx -- C ABI of the B200-native SPH / PCISPH time-step engine (drop-in for the simulation step of
 * felpzOliveira/Bubbles).
 *
 * Bubbles has no plugin / FFI layer: the seam this ABI replaces is the C++ solver API plus the
 * `*GPU` wrapper functions operating on managed-memory structs (reference paths are relative to the
 * Bubbles source tree):
 *     src/core/pcisph_solver.h:56-83   PciSphSolver3::{Initialize,Setup,SetColliders,Advance,...}
 *     src/core/sph_solver.h:76-96      SphSolver3::{...,Advance}
 *     src/core/sph_solver.h:171-195    UpdateGridDistributionGPU, ComputeDensityGPU, ... (the kernels)
 *     src/core/pcisph_solver.h:89-91   ComputePressureForceAndIntegrate
 * A C++ facade with the reference's class names lives in bubbles_b200/host/bubbles_api.h; the
 * binding a Bubbles maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions: POD only; every call returns an int status (0 = BBX_OK) and never exits the process
 * (the reference prints, waits on getchar() and exit(0)s: src/cuda/cutil.cpp:15-26); the engine owns
 * all device memory, the caller owns all host buffers; one host thread per engine handle; particle
 * ids are append order (as ParticleSetBuilder3::AddParticle, src/core/particle.h:631-636) and every
 * download is in original-id order.  There is NO CPU fallback: without a CUDA device bbx_create fails.
 */
#ifndef BBX_H
#define BBX_H

#ifdef __cplusplus
extern "C" {
#endif

#define BBX_VERSION 1
#define BBX_MAX_NEIGHBORS 100 /* MaximumParticlesPerBucket, src/core/particle.h:10 */
#define BBX_MAX_COLLIDERS 16

enum bbx_status {
    BBX_OK = 0,
    BBX_ERR_INVALID = 1,      /* bad argument / call order                                   */
    BBX_ERR_CUDA = 2,         /* CUDA runtime error (text in bbx_last_error)                 */
    BBX_ERR_NO_DEVICE = 3,    /* no usable GPU: the engine has no CPU path                   */
    BBX_ERR_CAPACITY = 4,     /* more particles than max_particles                           */
    BBX_ERR_OUT_OF_DOMAIN = 5,/* a particle position is outside the grid bounds              */
    BBX_ERR_COMM = 6          /* NCCL error, or a slab neighbour that never signalled its halo */
};

enum bbx_solver { BBX_SOLVER_PCISPH = 0, BBX_SOLVER_SPH = 1 };

/* Fields for bbx_download (original-id order).  vec3 fields are 3 values per particle. */
enum bbx_field {
    BBX_POSITION = 0,       /* ParticleSet3::positions   (src/core/particle.h:158)           */
    BBX_VELOCITY = 1,       /* ParticleSet3::velocities                                      */
    BBX_FORCE = 2,          /* ParticleSet3::forces (total force of the last sub-step)       */
    BBX_DENSITY = 3,        /* ParticleSet3::densities                                       */
    BBX_PRESSURE = 4,       /* ParticleSet3::pressures                                       */
    BBX_PRED_POSITION = 5,  /* PciSphSolverData3::tempPositions (src/core/pcisph_solver.h:19)*/
    BBX_PRED_DENSITY = 6,   /* PciSphSolverData3::densityPredicted                           */
    BBX_PRESSURE_FORCE = 7, /* PciSphSolverData3::pressureForces                             */
    BBX_FORCE_NP = 8,       /* ParticleSet3::forces after ComputeNonPressureForce            */
    BBX_DENSITY_ERROR = 9,  /* rho* - rho0 (what PredictPressureFor computes but never stores)*/
    BBX_NEIGHBOR_COUNT = 10 /* Bucket::Count() (int)                                         */
};

enum bbx_dtype { BBX_F32 = 0, BBX_F64 = 1, BBX_I32 = 2 };

/* Phases of one PCISPH sub-step, for phase-by-phase parity checks (bbx_run_phase).  Running
 * GRID..INTEGRATE in order equals bbx_step_pcisph in reference-compat mode. */
enum bbx_phase {
    BBX_PHASE_GRID = 0,        /* UpdateGridDistributionGPU (sph_equations3.cpp:511-539) incl. neighbour lists */
    BBX_PHASE_DENSITY = 1,     /* ComputeDensityGPU (sph_equations3.cpp:549-553)                               */
    BBX_PHASE_FORCE_NP = 2,    /* ComputeParticleInteractionGPU + ComputeNonPressureForceGPU (:593-620)        */
    BBX_PHASE_PREDICT = 3,     /* PredictVelocityAndPositionGPU (pcisph_equations3.cpp:50-56), first iteration */
    BBX_PHASE_PRESSURE = 4,    /* PredictPressureGPU (:108-112)                                                */
    BBX_PHASE_PRESSURE_FORCE = 5, /* PredictPressureForceGPU (:173-177)                                        */
    BBX_PHASE_INTEGRATE = 6    /* AccumulateAndIntegrateGPU (:207-212) + pseudo-viscosity (cold)               */
};

/* BBX_COLLIDER_MESH = a Shape of type ShapeMesh (MakeMesh, src/shapes/bvh.cpp:51-56) with the SDF grid the collider set
 * generates for it (ColliderSet3::GenerateSDFs -> GenerateShapeSDF, src/core/shape.cpp:479-511): the triangles decide which
 * collider is nearest (Shape::MeshClosestDistance through a BVH, and only for points inside the mesh bounds:
 * Collider3::OptmizedClosestPointCheck, collider.cpp:113-121), the grid gives closest point, normal and inside test. */
enum bbx_collider_type { BBX_COLLIDER_BOX = 0, BBX_COLLIDER_SPHERE = 1, BBX_COLLIDER_SDF = 2, BBX_COLLIDER_MESH = 3 };

/* One collider = Collider3 + Shape (src/core/collider.h:55-73, src/core/shape.h:140-170).
 * Matrices are row-major 4x4 (Transform::m / mInv, src/core/transform.h:399-416). */
typedef struct bbx_collider {
    int type;                 /* bbx_collider_type                                           */
    int reverse_orientation;  /* Shape::reverseOrientation (containers)                      */
    int active;               /* Collider3::isActive                                         */
    int reserved;
    double object_to_world[16];
    double world_to_object[16];
    double size[3];           /* box: sizex, sizey, sizez                                    */
    double radius;            /* sphere                                                      */
    double friction;          /* Collider3::frictionCoefficient                              */
    double linear_velocity[3];
    double angular_velocity[3];
    /* SDF field grid (vertex centred FieldGrid3f, src/core/grid.h:1218-1240): node counts,
     * node spacing, position of node (0,0,0), field values x-fastest (LinearIndex) */
    int sdf_resolution[3];
    int reserved2;
    double sdf_spacing[3];
    double sdf_origin[3];
    const double *sdf_field;  /* host pointer, copied by bbx_set_colliders                   */
    /* triangle mesh (BBX_COLLIDER_MESH; world space, as Transform::Mesh leaves it): copied by bbx_set_colliders */
    int mesh_vertices, mesh_triangles;
    const double *mesh_points;   /* mesh_vertices x 3                                        */
    const int *mesh_indices;     /* mesh_triangles x 3                                       */
} bbx_collider;

/* Grid geometry = result of UtilBuildGridForDomain / MakeGrid (src/core/util.cpp:269-287,
 * src/core/grid.h:626-666) */
typedef struct bbx_grid_desc {
    double min[3], max[3], cell_len[3];
    int n[3];
    int total;
} bbx_grid_desc;

/* All constants of SphSolverData3 + PciSphSolver3 (src/core/sph_solver.h:31-50,
 * src/solvers/sph_solver3.cpp:124-151, src/solvers/pcisph_solver3.cpp:9-14) */
typedef struct bbx_config {
    int struct_size;          /* sizeof(bbx_config), ABI check                               */
    int device;               /* CUDA device ordinal                                         */
    int max_particles;        /* capacity (reserved size of the particle set)                */
    int pcisph_max_iterations;/* 5                                                           */
    int pcisph_reference_compat; /* 1: one predict-correct iteration like the reference's
                                    effective behaviour (density error never stored);
                                    0: store the error and iterate to tolerance              */
    int with_gravity;         /* only used by bbx_config_default                             */
    double spacing;           /* target spacing = particle radius (src/core/particle.h:538)  */
    double kernel_scale;      /* kernelRadiusOverTargetSpacing                               */
    double target_density;    /* WaterDensity = 1000                                         */
    double viscosity;         /* 0.04                                                        */
    double drag;              /* 1e-4                                                        */
    double eos_exponent;      /* 7                                                           */
    double sound_speed;       /* 100                                                         */
    double negative_pressure_scale; /* 0                                                     */
    double pseudo_viscosity;  /* 10                                                          */
    double gravity[3];        /* sum of constant interactions: (0, -9.8f, 0)                 */
    double pcisph_max_density_error_ratio; /* 0.01                                           */
    double restitution;       /* 0.6 in TimeIntegrationFor (sph_equations3.cpp:307)          */
    double time_step_limit_scale; /* kDefaultTimeStepLimitScale = 5 (pcisph_solver2.cpp:7)   */
    bbx_grid_desc grid;       /* domain grid                                                 */
    /* multi-GPU slab decomposition: this engine owns the global cell planes [slab_z_begin, slab_z_end)
     * of `grid` (which always describes the WHOLE domain).  0, 0 (or 0, grid.n[2]) = single domain.      */
    int slab_z_begin, slab_z_end;
    int ghost_capacity;       /* slab engines: particle slots per ghost plane (0 = max_particles / 4)  */
    int reserved;
} bbx_config;

typedef struct bbx_step_stats {
    int particles;            /* owned particle count                                        */
    int ghosts;               /* ghost particles (multi-GPU)                                 */
    int substeps;             /* sub-steps done since creation                               */
    int pcisph_iterations;    /* predict-correct iterations of the last sub-step             */
    int full_rebuild;         /* last grid update was a full (ascending-id) rebuild          */
    int rebuild_flag;         /* SphParticleSet3::requiresHigherLevelUpdate after last step  */
    int neighbor_overflow;    /* particles whose list hit the 100 cap in the last update     */
    int lost_particles;       /* particles that moved >= 2 cells in an incremental update    */
    int clamped;              /* particles pushed back into the domain in the last sub-step  */
    int nan_count;            /* non-finite positions detected in the last sub-step          */
    float max_force;          /* max |f| after the last sub-step (CFL scan, particle.h:591)  */
    float max_density_error;  /* max |rho* - rho0| of the last iteration                     */
    float ms_grid;            /* device ms of the last sub-step's grid phase (when timing on)*/
    float ms_step;            /* device ms of the last sub-step                              */
    int exact_passes;         /* list-build passes redone with the FP64 IsWithinStd predicate  */
    int max_candidates;       /* largest 27-cell neighbourhood of the last list build          */
    int occupied_cells;       /* occupied (owned) cells of the last grid update                */
    int unstaged_tiles;       /* sweep tiles of the last sub-step whose neighbourhood did not fit in shared memory */
} bbx_step_stats;

typedef struct bbx_engine bbx_engine;

/* -- setup ------------------------------------------------------------------------------------ */
const char *bbx_last_error(void);
int bbx_version(void);
/* DefaultSphSolverData3(with_gravity) + PciSphSolver3::Initialize defaults */
int bbx_config_default(bbx_config *cfg, int with_gravity);
/* UtilBuildGridForDomain(Bounds3f, spacing, spacingScale) -- host arithmetic only */
int bbx_grid_for_domain(const double domain_min[3], const double domain_max[3], double spacing,
                        double kernel_scale, bbx_grid_desc *out);
/* MakeGrid(resolution, p0, p1) */
int bbx_grid_build(const int resolution[3], const double p0[3], const double p1[3], bbx_grid_desc *out);
/* PciSphSolver3::Setup / SphSolver3::Setup: allocates device state, computes mass and delta denom */
int bbx_create(const bbx_config *cfg, bbx_engine **out);
int bbx_destroy(bbx_engine *e);
/* Solver constants that the reference reads live from SphSolverData3 / PciSphSolver3 every sub-step
 * (SphSolver3::SetViscosityCoefficient, SetPseudoViscosityCoefficient, src/solvers/sph_solver3.cpp:22-33): change
 * one after bbx_create; takes effect with the next sub-step. */
enum bbx_param {
    BBX_PARAM_VISCOSITY = 0, BBX_PARAM_PSEUDO_VISCOSITY = 1, BBX_PARAM_DRAG = 2, BBX_PARAM_RESTITUTION = 3,
    BBX_PARAM_NEGATIVE_PRESSURE_SCALE = 4, BBX_PARAM_REFERENCE_COMPAT = 5, BBX_PARAM_MAX_ITERATIONS = 6,
    BBX_PARAM_MAX_DENSITY_ERROR_RATIO = 7, BBX_PARAM_TIME_STEP_LIMIT_SCALE = 8,
    BBX_PARAM_GRAVITY_X = 9, BBX_PARAM_GRAVITY_Y = 10, BBX_PARAM_GRAVITY_Z = 11
};
int bbx_set_param(bbx_engine *e, int param, double value);
/* scalars computed at setup: ParticleSet3::GetMass, PciSphSolver3::deltaDenom, ::ComputeDelta(dt) */
int bbx_get_mass(bbx_engine *e, double *mass);
int bbx_get_delta(bbx_engine *e, double dt, double *delta);

/* -- particles --------------------------------------------------------------------------------- */
/* SphParticleSet3FromBuilder + Setup's initial DistributeByParticle: replaces all particles;
 * pos / vel are n x 3 (AoS), dtype BBX_F32 or BBX_F64 */
int bbx_set_particles(bbx_engine *e, int n, const void *pos, const void *vel, int dtype);
/* slab engines: same, with explicit global particle ids (NULL = the index in the arrays).  A slab engine
 * keeps the particles whose cell plane it owns and ignores the rest, so every rank may pass the whole
 * scene or just its share.  Collective over the slab group (ends with the first ghost exchange).        */
int bbx_set_particles_ids(bbx_engine *e, int n, const void *pos, const void *vel, const int *ids, int dtype);
/* ContinuousParticleSetBuilder3::AddParticle + Commit (grid.h:1409-1441): new ids continue from the current
 * count, the new particles go to the tail of their cell's chain (DistributeByParticleList, grid.h:358-387), the
 * chains of the existing particles are left as they are.  Works before and between steps (neighbour lists are
 * refreshed by the next sub-step); slab engines: bbx_append_particles_ids. */
int bbx_append_particles(bbx_engine *e, int n, const void *pos, const void *vel, int dtype);
/* the same with explicit global ids, which is how particles are appended to SLAB engines (collective over the group):
 * every rank passes the appended particles (or any superset of its share); ids must exceed every id already in the run.
 * Single-domain engines accept ids == NULL or the continuing sequence. */
int bbx_append_particles_ids(bbx_engine *e, int n, const void *pos, const void *vel, const int *ids, int dtype);
int bbx_particle_count(bbx_engine *e, int *n);
/* overwrite positions+velocities of the existing particles (id order) without touching chains */
int bbx_overwrite_state(bbx_engine *e, const void *pos, const void *vel, int dtype);
/* slab engines (works on any engine): the same for the OWNED particles, rows in the order of the last
 * bbx_download_owned (the engine's cell order); the neighbours' ghost copies of the boundary planes are
 * refreshed, so the call is collective over the slab group.                                            */
int bbx_overwrite_owned(bbx_engine *e, const void *pos, const void *vel, int dtype);

/* -- colliders --------------------------------------------------------------------------------- */
int bbx_set_colliders(bbx_engine *e, int n, const bbx_collider *colliders);
int bbx_update_collider(bbx_engine *e, int index, const bbx_collider *collider); /* Shape::Update/SetVelocities */
int bbx_set_collider_active(bbx_engine *e, int index, int active);             /* ColliderSet3::SetActive   */
/* Shape::ClosestDistance of collider `index` (src/core/shape.cpp:206-233) at n points (x, y, z FP64): box / sphere signed,
 * SDF grid sample, mesh = distance to the nearest triangle through the device BVH (parity tests) */
int bbx_collider_distance(bbx_engine *e, int index, int n, const double *points, double *out);

/* -- stepping ---------------------------------------------------------------------------------- */
/* AdvanceTimeStep(PciSphSolver3*, dt) (src/solvers/pcisph_solver3.cpp:42-65): one sub-step */
int bbx_step_pcisph(bbx_engine *e, double dt);
/* AdvanceTimeStep(SphSolver3*, dt) (src/solvers/sph_solver3.cpp:48-67), Jacobi semantics */
int bbx_step_sph(bbx_engine *e, double dt);
/* PciSphSolver3::Advance / SphSolver3::Advance: CFL sub-stepping (particle.h:584-608) */
int bbx_advance(bbx_engine *e, double seconds, int solver, int *substeps, float *ms);
/* n fixed-dt sub-steps enqueued back to back without host synchronisation */
int bbx_step_many(bbx_engine *e, double dt, int solver, int n);
/* the same, bracketed by two events on the engine's stream: *ms = device time of the n sub-steps, launch gaps
 * included (what bench.py reports); returns once the last sub-step has finished */
int bbx_step_many_timed(bbx_engine *e, double dt, int solver, int n, float *ms);
int bbx_run_phase(bbx_engine *e, int phase, double dt);
int bbx_synchronize(bbx_engine *e);
/* Device-side failures (a particle outside the grid, a slab neighbour that never signalled, more particles than
 * slots) are sticky: the sub-step that hits one keeps running on what it has, the NEXT stepping call -- or any
 * synchronising call (bbx_synchronize, bbx_step_many_timed, bbx_advance, bbx_stats) -- returns the code, and
 * bbx_set_particles clears it.  The reference aborts the process in these situations (AssertA / exit). */
/* device-event timing of every sub-step phase (off by default; costs a sync per sub-step) */
int bbx_set_timing(bbx_engine *e, int enabled);
int bbx_stats(bbx_engine *e, bbx_step_stats *out);

/* -- results ----------------------------------------------------------------------------------- */
int bbx_download(bbx_engine *e, int field, void *dst, int dtype);
/* slab engines (works on any engine): the owned particles in the engine's cell order; ids[k] = global id
 * of row k of dst, *count = owned particles.  dst or ids may be NULL.                                   */
int bbx_download_owned(bbx_engine *e, int field, void *dst, int dtype, int *ids, int *count);
/* positions and velocities of one frame in a single call (one kernel, one synchronisation): rows in particle-id
 * order (owned_order = 0, single-domain engines) or in the engine's cell order with ids[k] = global id of row k
 * (owned_order = 1, any engine; ids may be NULL).  *count = rows written.  What UtilRunSimulation3 reads after every
 * Advance (src/core/util.h:583-590). */
int bbx_download_state(bbx_engine *e, void *pos, void *vel, int *ids, int dtype, int owned_order, int *count);
/* active chains: cell_count[total], cell_order[n] (concatenated chains, original ids) */
int bbx_export_cells(bbx_engine *e, int *cell_count, int *cell_order);
/* stored neighbour lists in the reference's bucket order: counts[n], ids[n*100] (unused = -1) */
int bbx_export_neighbors(bbx_engine *e, int *counts, int *ids);
/* same for the owned particles of a slab engine, rows in the order of bbx_download_owned (ids are global) */
int bbx_export_neighbors_owned(bbx_engine *e, int *counts, int *ids);
/* The per-cell test of ContinuousParticleSetBuilder3::MapGridEmit (src/core/grid.h:1367-1407) on the device, for a batch
 * of n template points (FP64 x, y, z) with the GLOBAL ids of their cells: cell_size[k] = length of the cell's current
 * chain (-1: a cell this slab engine does not own), blocked[k] = 1 when a particle of that chain lies closer than d
 * (the reference walks the chains on the host; this keeps them on the device -- only 8 bytes per point come back). */
int bbx_query_cells(bbx_engine *e, int n, const int *cells, const double *points, double d, int *cell_size, int *blocked);
/* replace the chain order (parity tests: reproduce a history-dependent order) */
int bbx_inject_chains(bbx_engine *e, const int *cell_count, const int *cell_order);
/* set SphParticleSet3::requiresHigherLevelUpdate */
int bbx_set_rebuild_flag(bbx_engine *e, int flag);
/* number of this library's kernels launched since creation (bench.py's gpu_launches) */
int bbx_launch_count(bbx_engine *e, long long *count);
/* device ms spent in each kernel family since the last reset (needs bbx_set_timing(1)) */
int bbx_kernel_time(bbx_engine *e, int phase, float *ms, int *launches);
int bbx_reset_kernel_time(bbx_engine *e);

/* -- multi-GPU (z-slab decomposition, NCCL over NVLink) ------------------------------------------
 * The reference is single-GPU; this part of the ABI is new.  One engine per GPU / process
 * owns a contiguous range of cell planes (bbx_config.slab_z_begin/end) plus one ghost plane per neighbour.
 * After bbx_comm_init every stepping call is COLLECTIVE over the group (all ranks call it with the same
 * arguments): the engine exchanges migrating particles, ghost planes (in the reference's chain order) and
 * the per-phase ghost fields with rank-1 / rank+1 through ncclSend/ncclRecv, and the global flags (big-move
 * rule, CFL force maximum, density error) through ncclAllReduce.  Cell orderings and neighbour lists of
 * the slabs concatenate to exactly the single-domain ones. */
#define BBX_NCCL_ID_BYTES 128
int bbx_comm_unique_id(unsigned char id[BBX_NCCL_ID_BYTES]);
int bbx_comm_init(bbx_engine *e, int rank, int nranks, const unsigned char id[BBX_NCCL_ID_BYTES]);
/* several slab engines of one process on one device, each driven by its own host thread (same code path,
 * copies instead of NCCL): lets the slab logic be verified against the single-domain engine on one GPU */
/* 1 when the sweeps store their boundary-plane results straight into the neighbours' ghost slots (peer
 * memory over NVLink, CUDA IPC), 0 when every phase ends with a send / recv pair (BBX_P2P=0, or no peer access) */
int bbx_halo_mode(bbx_engine *e, int *p2p);
int bbx_comm_init_local(bbx_engine *e, int rank, int nranks, const char *group);
/* plane_counts[nplanes] -> z_bounds[nranks + 1]: slabs of whole planes balanced by particle count */
int bbx_slab_plan(int nplanes, const long long *plane_counts, int nranks, int *z_bounds);
/* Re-balancing during a run (the reference has no decomposition; the static plan of bbx_slab_plan drifts as the fluid
 * moves).  bbx_plane_counts: owned particles per GLOBAL cell plane (grid n[2] entries, zero outside this engine's planes;
 * sum them over the group and feed bbx_slab_plan).  bbx_rebalance: collective over the group, between two sub-steps;
 * every rank passes the same z_bounds[nranks + 1].  Whole planes change hands between neighbouring ranks only (a rank
 * must keep at least one of its planes: move far cuts in several calls), in their chain order: results stay
 * bit-identical to the single-domain engine. */
int bbx_plane_counts(bbx_engine *e, long long *plane_counts);
int bbx_rebalance(bbx_engine *e, const int *z_bounds);
/* current[nranks + 1], target[nranks + 1] -> step[nranks + 1]: the part of the way to `target` one bbx_rebalance can go
 * (every cut stays strictly inside the two slabs it separates); *done = 1 when step == target */
int bbx_slab_plan_step(int nranks, const int *current, const int *target, int *step, int *done);
/* particles per global cell plane (host arithmetic, same hash as the engine) */
int bbx_plane_histogram(const bbx_grid_desc *grid, int n, const void *pos, int dtype, long long *plane_counts);

#ifdef __cplusplus
}
#endif
#endif /* BBX_H */
