// bbx engine: host side of the C ABI declared in include/bbx.h.  Owns all device state, enqueues
// the kernels of bbx_kernels.cuh on one CUDA stream, never exits the process, has no CPU path.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <vector>
#include <string>
#include <algorithm>
#include <mutex>

#include "../../include/bbx.h"
#include "bbx_kernels.cuh"
#include "bbx_host_math.h"
#include "bbx_comm.h"

static thread_local std::string g_last_error;
static int set_error(int code, const char *fmt, ...){
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    g_last_error = buf;
    return code;
}
#define CU(call) do{ cudaError_t _e = (call); if(_e != cudaSuccess) return set_error(BBX_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); }while(0)
#define CHECK_ENGINE(e) do{ if(!(e)) return set_error(BBX_ERR_INVALID, "null engine"); CU(cudaSetDevice((e)->device)); }while(0)

#define BBX_HINT_RING 4
#define BBX_HINT_LAG 2
enum { T_GRID = 0, T_DENSITY, T_FORCE_NP, T_PREDICT, T_PRESSURE, T_PRESSURE_FORCE, T_INTEGRATE, T_COUNT };

struct bbx_engine {
    bbx_config cfg;
    int device;
    cudaStream_t stream;
    cudaStream_t side; cudaEvent_t ev_fork, ev_join; // slab engines: the global flag reduction runs beside the scan and the fill
    int n, cap;      // owned particles / capacity
    // slab engines: ghost capacity per side and the current ghost / boundary-plane counts (host copies).
    // Every slot-indexed array that neighbours are read from (pos, vel, pid, cell, newcell, pred, posq) is
    // allocated with gcap slots in front: pointers are offset so that owned slots are [0, n), the lower
    // ghost plane is [-n_glo, 0) and the upper one [n, n + n_ghi).
    int gcap, n_glo, n_ghi, n_first, n_last;
    int has_lo, has_hi; // a slab neighbour exists below / above
    BbxComm *comm;
    // halo push: the neighbours' arrays as seen from this device (slot-0 pointers), their owned counts, and the
    // flags the neighbours raise in MY memory when their stores into my ghost slots are complete
    struct PeerSide { float4 *pos[2], *vel[2], *rec, *pred, *posq; int *pid[2], *gtab; unsigned *flags; int n; long long gc; } peer[2]; // [0] lower, [1] upper neighbour (gc = ITS ghost capacity per side)
    int peers_ready;    // share_arrays done (lazily, at the first collective grid update)
    int p2p;            // boundary-plane results are stored straight into the neighbours' ghost slots (else: send / recv per phase)
    unsigned *halo_flags;            // device, [2 sides][BBX_HALO_PHASES]: [0][p] raised by the lower neighbour, [1][p] by the upper one;
                                     // then [2 sides][BBX_HALO_MAIL] mailbox integers written by the neighbours
    int *mail_host;                  // pinned copy of the mailbox
    unsigned halo_seq[BBX_HALO_PHASES];
    void *raw_shared[BBX_PEER_NPTR]; // allocation bases of pos[2], vel[2], rec, pred, posq, halo_flags, pid[2], gtab
    int *gtab;          // received cell-table slices of the two ghost planes, 2 x (plane + 1)
    std::vector<void *> raw; // allocation bases (for cudaFree)
    DevGrid grid;
    double mass, delta_denom, mass_over_rho0_sq, h;
    // sorted particle arrays, double buffered for the reorder
    float4 *pos[2], *vel[2];
    int *pid[2], *cell[2];
    int *cell_start[2];
    int cur;
    int have_chains; // cell_start[cur] / cell[cur] valid
    int *newcell, *count, *perm, *occ_cells, *queue;
    unsigned *movemask; // per old cell: directions its leaving particles took (k_hash_count -> k_fill_incremental)
    unsigned long long *scan_status;
    int scan_tiles;
    int epoch;       // grid updates done so far; flags are double-buffered by its parity (DevState)
    int masked;      // cells smaller than h (informational; the cell-centric list build needs no window test)
    int list_ctas_per_sm;
    long long cells_alloc; // entries of the cell-indexed arrays (slab engines: enough for any slab of the grid)
    int *queue_b;          // collider slow-path queue of the boundary passes (slab engines with the halo push)
    int overlap;           // halo push: boundary blocks first, their halo travels while the interior blocks run (BBX_OVERLAP=0: off)
    unsigned halo_due[BBX_HALO_PHASES]; // sequence number of a halo this engine has signalled but not yet waited for (0: none)
    cudaStream_t bstream;  // the boundary blocks of a phase run here, beside the interior blocks on `stream`
    cudaEvent_t ev_b[4], ev_i[4], ev_grid; // boundary / interior blocks of phase k are done; the grid update is done
    unsigned short *nbr; int *nbr_cnt;
    float4 *force, *force_p, *pred, *posq, *smoothed;
    float4 *rec;     // 32-byte gather records (x, y, z, rho | vx, vy, vz, -), 2 float4 per slot, written by the list build
    float *pressure, *rho_pred, *rho_err;
    DevState *st; DevState *st_host; // st_host pinned
    int *err_probe;  // pinned: copy of st->error enqueued behind every sub-step (no sync); checked by the next API call
    int sm_count;    // multiprocessors of the device (persistent grids are sized from it)
    // slab engines with the halo push launch their per-particle kernels over a BOUND of the owned count: the count of the grid
    // update BBX_HINT_LAG sub-steps back (an asynchronous copy of DevState behind every update, in a small ring; the host
    // waits for THAT old copy only, which also keeps it from running more than a few sub-steps ahead of the device) plus a
    // margin; k_slab_plan checks the real count against the bound on the device (sticky capacity error if ever too small)
    DevState *st_hint; cudaEvent_t ev_hint[BBX_HINT_RING]; long long hint_count; int n_hint; int n_launch;
    int counts_stale;   // slab engines with the halo push: n / n_first / n_last / ghost counts live in DevState; the host copies
                        // (e->n ...) are refreshed by sync_counts() when an API call needs them
    int deferred_error; // slab engines: a device-side error seen in the middle of a collective sub-step; the sub-step is
                        // finished (so that the neighbours are not left waiting for this rank's halo signals), then returned
    DevColliderSet *colliders; DevColliderSet colliders_host;
    DevCullSet *cull; DevCullSet cull_host;
    std::vector<double *> sdf_fields;
    std::vector<float *> sdf_fields32;
    std::vector<void *> mesh_allocs; // device copies of mesh colliders: vertices, triangles in BVH leaf order, BVH nodes
    void *stage; size_t stage_bytes; // device staging for upload / download
    long long launches;
    int substeps;
    int timing;
    std::vector<cudaEvent_t> ev; size_t ev_used;
    std::vector<int> ev_phase;
    float phase_ms[T_COUNT + 1]; int phase_launches[T_COUNT + 1]; // [T_COUNT] = gaps between sub-steps
    float last_ms_grid, last_ms_step;
    int force_full; // next grid update must be a full rebuild (fresh particle set)
    int last_force; // the last grid update was forced to the full path by the host
};

#define BBX_PAD 64 // spare slots at the end of the particle arrays
static int push_cull(bbx_engine *e);
#define LAUNCH(e, kernel, grid, block, ...) do{ kernel<<<(grid), (block), 0, (e)->stream>>>(__VA_ARGS__); (e)->launches++; }while(0)
#define LAUNCH_S(e, kernel, grid, block, smem, ...) do{ kernel<<<(grid), (block), (smem), (e)->stream>>>(__VA_ARGS__); (e)->launches++; }while(0)
#define LAUNCH_ON(e, strm, kernel, grid, block, smem, ...) do{ kernel<<<(grid), (block), (smem), (strm)>>>(__VA_ARGS__); (e)->launches++; }while(0)
static inline int div_up(long long a, int b){ return (int)((a + b - 1) / b); }

static const char *device_error_text(int code){
    return code == BBX_ERR_OUT_OF_DOMAIN ? "particle outside the grid bounds" :
           (code == BBX_ERR_COMM ? "a slab neighbour never signalled its halo stores" :
           (code == BBX_ERR_CAPACITY ? "more particles than slots (max_particles / ghost_capacity), or a cell run longer than 4096 particles" : "device-side error"));
}
// Device-side errors are sticky in DevState::error.  Every sub-step ends with an asynchronous copy of that word into
// pinned host memory (no synchronisation); the stepping calls look at it on entry, the synchronising calls after
// their sync -- so a failure surfaces at the latest one call after the sub-step that hit it, and
// bbx_set_particles clears it.
static int sticky_error(bbx_engine *e){
    const int code = e->err_probe ? *(volatile int *)e->err_probe : 0;
    if(code) return set_error(code, "device-side error %d (%s)", code, device_error_text(code));
    return BBX_OK;
}
// entry of a sub-step.  On a slab engine a particle that left the halo (BBX_ERR_OUT_OF_DOMAIN: dropped by the hash kernel,
// the state stays consistent) must not make this rank skip a collective sub-step its neighbours are about to wait in:
// the sub-step runs and the code is returned on the way out (deferred_error).
static int entry_error(bbx_engine *e){
    const int code = e->err_probe ? *(volatile int *)e->err_probe : 0;
    if(code == BBX_ERR_OUT_OF_DOMAIN && (e->has_lo || e->has_hi)){ e->deferred_error = code; return BBX_OK; }
    return sticky_error(e);
}
static void probe_error(bbx_engine *e){
    if(e->err_probe) cudaMemcpyAsync(e->err_probe, &e->st->error, sizeof(int), cudaMemcpyDeviceToHost, e->stream);
}

const char *bbx_last_error(void){ return g_last_error.c_str(); }
int bbx_version(void){ return BBX_VERSION; }

int bbx_config_default(bbx_config *cfg, int with_gravity){
    if(!cfg) return set_error(BBX_ERR_INVALID, "null config");
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = (int)sizeof(bbx_config);
    cfg->pcisph_max_iterations = 5;
    cfg->pcisph_reference_compat = 1;
    cfg->with_gravity = with_gravity;
    cfg->spacing = 0.1; cfg->kernel_scale = 2.0; cfg->target_density = 1000.0;
    cfg->viscosity = 0.04; cfg->drag = 0.0001; cfg->eos_exponent = 7.0; cfg->sound_speed = 100.0;
    cfg->negative_pressure_scale = 0.0; cfg->pseudo_viscosity = 10.0;
    if(with_gravity){ cfg->gravity[0] = 0.f; cfg->gravity[1] = -9.8f; cfg->gravity[2] = 0.f; }
    cfg->pcisph_max_density_error_ratio = 0.01;
    cfg->restitution = 0.6;
    cfg->time_step_limit_scale = 5.0;
    return BBX_OK;
}

int bbx_grid_for_domain(const double dmin[3], const double dmax[3], double spacing, double scale, bbx_grid_desc *out){
    if(!dmin || !dmax || !out || !(spacing > 0) || !(scale > 0)) return set_error(BBX_ERR_INVALID, "bad grid arguments");
    bbxh_grid_for_domain(dmin, dmax, spacing, scale, out);
    return BBX_OK;
}
int bbx_grid_build(const int res[3], const double p0[3], const double p1[3], bbx_grid_desc *out){
    if(!res || !p0 || !p1 || !out) return set_error(BBX_ERR_INVALID, "bad grid arguments");
    bbxh_grid_build(res, p0, p1, out);
    return BBX_OK;
}

template<typename T> static int dev_alloc(T **p, size_t count){
    CU(cudaMalloc((void **)p, sizeof(T) * (count ? count : 1)));
    return BBX_OK;
}

static int create_engine(const bbx_config *cfg, bbx_engine **slot);
// CUDA loads a kernel lazily at its first launch, and that load synchronises the context.  A first launch issued while a
// neighbour's k_halo_wait is spinning on this very rank's signal (several slab engines of ONE process on ONE device: the
// LocalComm test transport) would therefore wait for the spin, which waits for the launch: 30 s of stall, then BBX_ERR_COMM.
// Every kernel is loaded here, once per process, before any engine steps.
static int preload_kernels(){
    static std::mutex m; static bool done = false;
    std::lock_guard<std::mutex> lk(m);
    if(done) return BBX_OK;
    cudaFuncAttributes a;
#define BBX_PRELOAD(k) CU(cudaFuncGetAttributes(&a, k))
    BBX_PRELOAD(k_append_hash); BBX_PRELOAD(k_cell_lists_density<0>); BBX_PRELOAD(k_cell_lists_density<1>);
    BBX_PRELOAD(k_lists_density_tp<0>); BBX_PRELOAD(k_lists_density_tp<1>);
    BBX_PRELOAD(k_collide_integrate); BBX_PRELOAD(k_collide_predict); BBX_PRELOAD(k_collider_distance);
    BBX_PRELOAD(k_download); BBX_PRELOAD(k_download_state); BBX_PRELOAD(k_export_cells); BBX_PRELOAD(k_export_neighbors);
    BBX_PRELOAD(k_fill_incremental); BBX_PRELOAD(k_force_np_predict); BBX_PRELOAD(k_full_gather); BBX_PRELOAD(k_full_scatter);
    BBX_PRELOAD(k_full_sort_cells); BBX_PRELOAD(k_ghost_table); BBX_PRELOAD(k_halo_signal); BBX_PRELOAD(k_halo_wait);
    BBX_PRELOAD(k_hash_count); BBX_PRELOAD(k_inject_gather); BBX_PRELOAD(k_integrate); BBX_PRELOAD(k_overwrite);
    BBX_PRELOAD(k_predict_again); BBX_PRELOAD(k_pressure); BBX_PRELOAD(k_pressure_force<0>); BBX_PRELOAD(k_pressure_force<1>);
    BBX_PRELOAD(k_pseudo_aggregate); BBX_PRELOAD(k_pseudo_interpolate); BBX_PRELOAD(k_push_planes); BBX_PRELOAD(k_query_cells);
    BBX_PRELOAD(k_scan_cells); BBX_PRELOAD(k_slab_counts); BBX_PRELOAD(k_slab_plan); BBX_PRELOAD(k_slab_plan_host);
    BBX_PRELOAD(k_slot_of_id); BBX_PRELOAD(k_sph_forces); BBX_PRELOAD(k_upload); BBX_PRELOAD(k_upload_slab);
#undef BBX_PRELOAD
    done = true;
    return BBX_OK;
}
int bbx_create(const bbx_config *cfg, bbx_engine **out){
    if(!cfg || !out) return set_error(BBX_ERR_INVALID, "null argument");
    if(cfg->struct_size != (int)sizeof(bbx_config)) return set_error(BBX_ERR_INVALID, "bbx_config size mismatch (%d vs %d)", cfg->struct_size, (int)sizeof(bbx_config));
    if(cfg->max_particles <= 0 || cfg->grid.total <= 0 || !(cfg->spacing > 0) || !(cfg->kernel_scale > 0))
        return set_error(BBX_ERR_INVALID, "max_particles, grid, spacing and kernel_scale must be set");
    int ndev = 0;
    if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return set_error(BBX_ERR_NO_DEVICE, "no CUDA device: bbx has no CPU fallback");
    if(cfg->device < 0 || cfg->device >= ndev) return set_error(BBX_ERR_INVALID, "device %d out of range (%d devices)", cfg->device, ndev);
    CU(cudaSetDevice(cfg->device));
    bbx_engine *e = nullptr;
    int rc = create_engine(cfg, &e);
    if(rc != BBX_OK){ std::string keep = g_last_error; bbx_destroy(e); g_last_error = keep; return rc; } // one cleanup path; first error kept
    *out = e;
    return BBX_OK;
}

static int create_engine(const bbx_config *cfg, bbx_engine **slot){
    bbx_engine *e = new bbx_engine();
    *slot = e; // from here on the caller destroys it on failure
    e->stream = nullptr; e->side = nullptr; e->ev_fork = nullptr; e->ev_join = nullptr; e->comm = nullptr;
    e->bstream = nullptr; e->ev_grid = nullptr; for(int k = 0; k < 4; k++){ e->ev_b[k] = nullptr; e->ev_i[k] = nullptr; }
    for(int b = 0; b < 2; b++){ e->pos[b] = e->vel[b] = nullptr; e->pid[b] = e->cell[b] = e->cell_start[b] = nullptr; }
    e->newcell = e->count = e->perm = e->occ_cells = e->queue = e->queue_b = nullptr; e->overlap = 0; memset(e->halo_due, 0, sizeof(e->halo_due)); e->movemask = nullptr; e->scan_status = nullptr;
    e->nbr = nullptr; e->nbr_cnt = nullptr; e->force = e->force_p = e->pred = e->posq = e->smoothed = e->rec = nullptr;
    e->pressure = e->rho_pred = e->rho_err = nullptr; e->st = nullptr; e->st_host = nullptr; e->err_probe = nullptr;
    e->colliders = nullptr; e->cull = nullptr; e->gtab = nullptr; e->halo_flags = nullptr; e->mail_host = nullptr; e->stage = nullptr;
    e->st_hint = nullptr; for(int k = 0; k < BBX_HINT_RING; k++) e->ev_hint[k] = nullptr; e->hint_count = 0; e->n_hint = 0; e->n_launch = 0;
    memset(&e->cfg, 0, sizeof(e->cfg));
    e->cfg = *cfg;
    e->device = cfg->device;
    e->n = 0; e->cap = cfg->max_particles; e->cur = 0; e->have_chains = 0; e->launches = 0; e->substeps = 0; e->epoch = 0;
    e->timing = 0; e->ev_used = 0; e->force_full = 1; e->stage = nullptr; e->stage_bytes = 0; e->deferred_error = 0; e->counts_stale = 0;
    e->last_ms_grid = e->last_ms_step = 0.f;
    memset(e->phase_ms, 0, sizeof(e->phase_ms)); memset(e->phase_launches, 0, sizeof(e->phase_launches));
    CU(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&e->bstream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&e->ev_grid, cudaEventDisableTiming));
    for(int k = 0; k < 4; k++){ CU(cudaEventCreateWithFlags(&e->ev_b[k], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&e->ev_i[k], cudaEventDisableTiming)); }
    CU(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
    CU(cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, e->device));
    if(e->sm_count < 1) e->sm_count = 1;
    { int rc_ = preload_kernels(); if(rc_) return rc_; }
#ifdef BBX_LISTS_V7
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e->list_ctas_per_sm, k_cell_lists_density<0>, BBX_LT, 0));
#else
    // the thread-per-particle list build stages its tiles' candidates in dynamic shared memory (> 48 KB: opt in)
    CU(cudaFuncSetAttribute(k_lists_density_tp<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BBX_TP_SMEM));
    CU(cudaFuncSetAttribute(k_lists_density_tp<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BBX_TP_SMEM));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e->list_ctas_per_sm, k_lists_density_tp<0>, BBX_TP_WARPS * 32, BBX_TP_SMEM));
#endif
    if(e->list_ctas_per_sm < 1) e->list_ctas_per_sm = 1;
    // the staged sweeps carry their tile's neighbourhood in dynamic shared memory (> 48 KB: opt in)
    CU(cudaFuncSetAttribute(k_pressure, cudaFuncAttributeMaxDynamicSharedMemorySize, BBX_STAGE_BYTES(3)));
    // grid
    DevGrid &g = e->grid;
    for(int k = 0; k < 3; k++){
        g.min[k] = cfg->grid.min[k]; g.max[k] = cfg->grid.max[k]; g.len[k] = cfg->grid.cell_len[k];
        g.minf[k] = (float)g.min[k]; g.maxf[k] = (float)g.max[k]; g.lenf[k] = (float)g.len[k]; g.n[k] = cfg->grid.n[k];
    }
    g.plane = g.n[0] * g.n[1];
    if((long long)g.n[0] * g.n[1] * g.n[2] != cfg->grid.total) return set_error(BBX_ERR_INVALID, "grid.total != nx*ny*nz");
    // z-slab: owned global planes [zb, ze); one ghost plane towards each existing neighbour
    g.gnz = g.n[2];
    int zb = 0, ze = g.gnz;
    if(cfg->slab_z_end > cfg->slab_z_begin){
        zb = cfg->slab_z_begin; ze = cfg->slab_z_end;
        if(zb < 0 || ze > g.gnz) return set_error(BBX_ERR_INVALID, "slab [%d, %d) outside the grid's %d planes", zb, ze, g.gnz);
    }
    e->has_lo = zb > 0; e->has_hi = ze < g.gnz;
    g.zoff = zb - (e->has_lo ? 1 : 0);
    g.n[2] = (ze - zb) + (e->has_lo ? 1 : 0) + (e->has_hi ? 1 : 0);
    g.own_z0 = e->has_lo ? 1 : 0; g.own_z1 = g.own_z0 + (ze - zb);
    g.c_own0 = g.own_z0 * g.plane; g.c_own1 = g.own_z1 * g.plane;
    if((long long)g.plane * g.n[2] > 0x7fffffffLL) return set_error(BBX_ERR_INVALID, "more than 2^31 local cells");
    g.total = g.plane * g.n[2];
    e->gcap = 0; e->n_glo = e->n_ghi = e->n_first = e->n_last = 0; e->comm = nullptr; e->gtab = nullptr;
    e->p2p = 0; e->peers_ready = 0; e->halo_flags = nullptr; e->mail_host = nullptr; memset(e->peer, 0, sizeof(e->peer)); memset(e->halo_seq, 0, sizeof(e->halo_seq));
    memset(e->raw_shared, 0, sizeof(e->raw_shared));
    if(e->has_lo || e->has_hi){
        e->gcap = cfg->ghost_capacity > 0 ? cfg->ghost_capacity : std::max(4096, cfg->max_particles / 4);
    }
    // scalars of Setup: SetTargetDensity/Spacing/RelativeKernelRadius -> ComputeMass; deltaDenom
    e->h = cfg->kernel_scale * cfg->spacing;
    e->mass = bbxh_compute_mass(e->h, cfg->spacing, cfg->target_density);
    double r = e->mass / cfg->target_density;
    e->mass_over_rho0_sq = r * r;
    e->delta_denom = bbxh_delta_denom(e->h, cfg->spacing);
    // the list build tests candidates of cells two columns away only through the distance: exact as long
    // as a cell is not (measurably) shorter than h, otherwise the explicit window test is compiled in
    {
        double minlen = std::min(g.len[0], std::min(g.len[1], g.len[2]));
        e->masked = (minlen * minlen >= e->h * e->h - 0.5e-8) ? 0 : 1;
    }
    // cell-indexed arrays: a slab engine is sized for any slab of the grid (bbx_rebalance moves the cuts during a run)
    e->cells_alloc = (e->has_lo || e->has_hi) ? (long long)g.plane * (g.gnz + 2) : (long long)g.total;
    if(e->cells_alloc > 0x7fffffffLL) return set_error(BBX_ERR_INVALID, "more than 2^31 cells");
    // particle arrays carry BBX_PAD spare slots: the list build reads (and discards) up to 7 slots past a run
    const size_t gc = (size_t)e->gcap;
    size_t cap = (size_t)e->cap + BBX_PAD, capw = ((cap + 31) / 32) * 32, capg = cap + 2 * gc;
    int rc = BBX_OK;
#define BBX_TRY(call) do{ if(rc == BBX_OK) rc = (call); }while(0) /* keeps the FIRST error */
#define BBX_ALLOC_G(ptr) do{ BBX_TRY(dev_alloc(&(ptr), capg)); if(rc == BBX_OK){ e->raw.push_back((void *)(ptr)); CU(cudaMemset((ptr), 0, sizeof(*(ptr)) * capg)); (ptr) += gc; }else (ptr) = nullptr; }while(0)
    for(int b = 0; b < 2 && rc == BBX_OK; b++){
        BBX_ALLOC_G(e->pos[b]); BBX_ALLOC_G(e->vel[b]); BBX_ALLOC_G(e->pid[b]); BBX_ALLOC_G(e->cell[b]);
        BBX_TRY(dev_alloc(&e->cell_start[b], (size_t)e->cells_alloc + 1));
        if(rc == BBX_OK) CU(cudaMemset(e->cell_start[b], 0, sizeof(int) * ((size_t)e->cells_alloc + 1)));
    }
    BBX_ALLOC_G(e->newcell); BBX_ALLOC_G(e->pred); BBX_ALLOC_G(e->posq);
#undef BBX_ALLOC_G
    {   // records: 2 float4 per slot, ghost slots in front like the other slot-indexed arrays (32-byte aligned)
        float4 *raw = nullptr;
        BBX_TRY(dev_alloc(&raw, 2 * capg));
        if(rc == BBX_OK){ e->raw.push_back((void *)raw); CU(cudaMemset(raw, 0, sizeof(float4) * 2 * capg)); e->rec = raw + 2 * gc; }
    }
    if(rc == BBX_OK){
        e->raw_shared[0] = e->pos[0] - gc; e->raw_shared[1] = e->pos[1] - gc; e->raw_shared[2] = e->vel[0] - gc; e->raw_shared[3] = e->vel[1] - gc;
        e->raw_shared[4] = e->rec - 2 * gc; e->raw_shared[5] = e->pred - gc; e->raw_shared[6] = e->posq - gc;
        e->raw_shared[8] = e->pid[0] - gc; e->raw_shared[9] = e->pid[1] - gc;
    }
    if(e->gcap){
        const size_t words = (size_t)2 * (BBX_HALO_PHASES + BBX_HALO_MAIL);
        BBX_TRY(dev_alloc(&e->halo_flags, words));
        if(rc == BBX_OK){ CU(cudaMemset(e->halo_flags, 0, sizeof(unsigned) * words)); e->raw_shared[7] = e->halo_flags; }
        CU(cudaMallocHost((void **)&e->mail_host, sizeof(int) * 2 * BBX_HALO_MAIL));
    }
    BBX_TRY(dev_alloc(&e->count, (size_t)e->cells_alloc + 8)); BBX_TRY(dev_alloc(&e->perm, cap));
    BBX_TRY(dev_alloc(&e->occ_cells, (size_t)e->cells_alloc)); BBX_TRY(dev_alloc(&e->queue, cap));
    if(e->has_lo || e->has_hi) BBX_TRY(dev_alloc(&e->queue_b, cap));
    BBX_TRY(dev_alloc(&e->movemask, (size_t)e->cells_alloc));
    e->scan_tiles = div_up(g.c_own1 - g.c_own0, SCAN_TILE);
    BBX_TRY(dev_alloc(&e->scan_status, (size_t)div_up(e->cells_alloc, SCAN_TILE)));
    BBX_TRY(dev_alloc(&e->nbr, capw * BBX_NBR_CHUNKS * 8)); BBX_TRY(dev_alloc(&e->nbr_cnt, cap));
    BBX_TRY(dev_alloc(&e->force, cap)); BBX_TRY(dev_alloc(&e->force_p, cap));
    BBX_TRY(dev_alloc(&e->smoothed, cap));
    BBX_TRY(dev_alloc(&e->pressure, cap)); BBX_TRY(dev_alloc(&e->rho_pred, cap)); BBX_TRY(dev_alloc(&e->rho_err, cap));
    BBX_TRY(dev_alloc(&e->st, 1)); BBX_TRY(dev_alloc(&e->colliders, 1)); BBX_TRY(dev_alloc(&e->cull, 1));
    if(e->gcap){ BBX_TRY(dev_alloc(&e->gtab, 2 * ((size_t)g.plane + 1))); e->raw_shared[10] = e->gtab; }
    if(rc != BBX_OK){ return rc; }
    CU(cudaMemset(e->count, 0, sizeof(int) * ((size_t)e->cells_alloc + 8)));
    CU(cudaMallocHost((void **)&e->st_host, sizeof(DevState)));
    CU(cudaMallocHost((void **)&e->err_probe, sizeof(int)));
    CU(cudaMallocHost((void **)&e->st_hint, BBX_HINT_RING * sizeof(DevState)));
    memset(e->st_hint, 0, BBX_HINT_RING * sizeof(DevState));
    for(int k = 0; k < BBX_HINT_RING; k++) CU(cudaEventCreateWithFlags(&e->ev_hint[k], cudaEventDisableTiming));
    *e->err_probe = 0;
    memset(e->st_host, 0, sizeof(DevState));
    e->st_host->cap = e->cap; e->st_host->gcap = e->gcap;
    CU(cudaMemcpy(e->st, e->st_host, sizeof(DevState), cudaMemcpyHostToDevice));
    memset(&e->colliders_host, 0, sizeof(DevColliderSet));
    CU(cudaMemset(e->colliders, 0, sizeof(DevColliderSet)));
    { int rc2 = push_cull(e); if(rc2) return rc2; }
    CU(cudaMemset(e->force, 0, sizeof(float4) * cap)); CU(cudaMemset(e->force_p, 0, sizeof(float4) * cap));
    CU(cudaMemset(e->pressure, 0, sizeof(float) * cap)); CU(cudaMemset(e->rho_pred, 0, sizeof(float) * cap));
    CU(cudaMemset(e->rho_err, 0, sizeof(float) * cap));
    CU(cudaMemset(e->nbr_cnt, 0, sizeof(int) * cap));
    return BBX_OK;
}

int bbx_destroy(bbx_engine *e){
    if(!e) return BBX_OK;
    cudaSetDevice(e->device);
    if(e->stream) cudaStreamSynchronize(e->stream);
    if(e->comm){ delete e->comm; e->comm = nullptr; }
    for(void *p : e->raw) cudaFree(p);
    for(int b = 0; b < 2; b++) cudaFree(e->cell_start[b]);
    cudaFree(e->movemask); cudaFree(e->count); cudaFree(e->perm); cudaFree(e->occ_cells); cudaFree(e->queue); cudaFree(e->queue_b); cudaFree(e->scan_status);
    cudaFree(e->nbr); cudaFree(e->nbr_cnt); cudaFree(e->force); cudaFree(e->force_p);
    cudaFree(e->smoothed); cudaFree(e->pressure); cudaFree(e->rho_pred); cudaFree(e->rho_err);
    if(e->gtab) cudaFree(e->gtab);
    if(e->halo_flags) cudaFree(e->halo_flags);
    if(e->mail_host) cudaFreeHost(e->mail_host);
    cudaFree(e->st); cudaFree(e->colliders); cudaFree(e->cull); if(e->st_host) cudaFreeHost(e->st_host);
    if(e->err_probe) cudaFreeHost(e->err_probe);
    if(e->st_hint) cudaFreeHost(e->st_hint);
    for(int k = 0; k < BBX_HINT_RING; k++) if(e->ev_hint[k]) cudaEventDestroy(e->ev_hint[k]);
    for(double *f : e->sdf_fields) cudaFree(f);
    for(float *f : e->sdf_fields32) cudaFree(f);
    for(void *m : e->mesh_allocs) cudaFree(m);
    if(e->stage) cudaFree(e->stage);
    for(cudaEvent_t ev : e->ev) cudaEventDestroy(ev);
    if(e->ev_fork) cudaEventDestroy(e->ev_fork);
    if(e->ev_join) cudaEventDestroy(e->ev_join);
    if(e->bstream){ cudaStreamSynchronize(e->bstream); cudaStreamDestroy(e->bstream); }
    if(e->ev_grid) cudaEventDestroy(e->ev_grid);
    for(int k = 0; k < 4; k++){ if(e->ev_b[k]) cudaEventDestroy(e->ev_b[k]); if(e->ev_i[k]) cudaEventDestroy(e->ev_i[k]); }
    if(e->side) cudaStreamDestroy(e->side);
    if(e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return BBX_OK;
}

int bbx_get_mass(bbx_engine *e, double *mass){ if(!e || !mass) return set_error(BBX_ERR_INVALID, "null"); *mass = e->mass; return BBX_OK; }
int bbx_get_delta(bbx_engine *e, double dt, double *delta){
    if(!e || !delta) return set_error(BBX_ERR_INVALID, "null");
    *delta = bbxh_delta(e->mass_over_rho0_sq, e->delta_denom, dt);
    return BBX_OK;
}

int bbx_set_param(bbx_engine *e, int param, double value){
    if(!e) return set_error(BBX_ERR_INVALID, "null engine");
    if(!(value == value)) return set_error(BBX_ERR_INVALID, "NaN parameter");
    bbx_config &c = e->cfg; // make_params reads it at the start of every sub-step
    switch(param){
        case BBX_PARAM_VISCOSITY: c.viscosity = value > 0 ? value : 0; break; // SetViscosityCoefficient clamps at 0
        case BBX_PARAM_PSEUDO_VISCOSITY: c.pseudo_viscosity = value; break;
        case BBX_PARAM_DRAG: c.drag = value; break;
        case BBX_PARAM_RESTITUTION: c.restitution = value; break;
        case BBX_PARAM_NEGATIVE_PRESSURE_SCALE: c.negative_pressure_scale = value; break;
        case BBX_PARAM_REFERENCE_COMPAT: c.pcisph_reference_compat = value != 0 ? 1 : 0; break;
        case BBX_PARAM_MAX_ITERATIONS: if(value < 1) return set_error(BBX_ERR_INVALID, "at least one iteration"); c.pcisph_max_iterations = (int)value; break;
        case BBX_PARAM_MAX_DENSITY_ERROR_RATIO: c.pcisph_max_density_error_ratio = value; break;
        case BBX_PARAM_TIME_STEP_LIMIT_SCALE: if(!(value > 0)) return set_error(BBX_ERR_INVALID, "scale must be positive"); c.time_step_limit_scale = value; break;
        case BBX_PARAM_GRAVITY_X: c.gravity[0] = value; break;
        case BBX_PARAM_GRAVITY_Y: c.gravity[1] = value; break;
        case BBX_PARAM_GRAVITY_Z: c.gravity[2] = value; break;
        default: return set_error(BBX_ERR_INVALID, "unknown parameter %d", param);
    }
    return BBX_OK;
}

static int ensure_stage(bbx_engine *e, size_t bytes){
    if(e->stage_bytes >= bytes) return BBX_OK;
    if(e->stage){ CU(cudaFree(e->stage)); e->stage = nullptr; e->stage_bytes = 0; }
    CU(cudaMalloc(&e->stage, bytes));
    e->stage_bytes = bytes;
    return BBX_OK;
}

static int read_state(bbx_engine *e);
static int grid_update(bbx_engine *e);
#define IS_SLAB(e) ((e)->has_lo || (e)->has_hi)
#define COMM(call) do{ if((call)) return set_error(BBX_ERR_COMM, "%s", e->comm->err.c_str()); }while(0)

// Slab engines keep their particle counts on the device (DevState) and launch the per-particle kernels over the capacity;
// the host copies are brought up to date only when an API call needs exact numbers (downloads, exports, the send / recv
// transport of the cold paths).
static inline int launch_n(const bbx_engine *e){ return IS_SLAB(e) ? e->n_launch : e->n; }
static inline int bound_of(const bbx_engine *e, int n){ return (int)std::min<long long>(e->cap, (long long)n + n / 16 + 4096); }
// the owned count of the grid update BBX_HINT_LAG sub-steps back (by now that copy has long landed unless the host is
// running ahead -- then this is where it waits, a couple of sub-steps of device work still queued behind it)
static int poll_hint(bbx_engine *e){
    if(e->hint_count >= BBX_HINT_LAG){
        const int slot = (int)((e->hint_count - BBX_HINT_LAG) % BBX_HINT_RING);
        CU(cudaEventSynchronize(e->ev_hint[slot]));
        e->n_hint = e->st_hint[slot].n_own;
    }
    return BBX_OK;
}
static int sync_counts(bbx_engine *e){
    if(!e->counts_stale) return BBX_OK;
    int rc = read_state(e); if(rc) return rc;
    const DevState &s = *e->st_host;
    e->n = s.n_own; e->n_first = s.n_first; e->n_last = s.n_last; e->n_glo = s.n_glo; e->n_ghi = s.n_ghi;
    e->n_hint = s.n_own; e->hint_count = 0;
    e->peer[0].n = s.peer_n[0]; e->peer[1].n = s.peer_n[1];
    e->counts_stale = 0;
    if(s.error == BBX_ERR_CAPACITY) return set_error(BBX_ERR_CAPACITY, "slab capacity exceeded: %d owned (max_particles %d), ghost planes %d / %d (ghost_capacity %d)", s.n_own, e->cap, s.n_glo, s.n_ghi, e->gcap);
    return BBX_OK;
}

static int upload_particles(bbx_engine *e, int first, int n, const void *pos, const void *vel, const int *ids, int dtype){
    if(dtype != BBX_F32 && dtype != BBX_F64) return set_error(BBX_ERR_INVALID, "dtype must be BBX_F32 or BBX_F64");
    size_t esz = dtype == BBX_F64 ? 8 : 4;
    size_t bytes = esz * 3 * (size_t)n, idb = (IS_SLAB(e) && ids) ? sizeof(int) * (size_t)n : 0;
    int rc = ensure_stage(e, 2 * bytes + idb); if(rc) return rc;
    char *sp = (char *)e->stage, *sv = sp + bytes; int *si = idb ? (int *)(sv + bytes) : nullptr;
    CU(cudaMemcpyAsync(sp, pos, bytes, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(sv, vel, bytes, cudaMemcpyHostToDevice, e->stream));
    if(IS_SLAB(e)){
        // keep the particles of the owned planes (the caller may pass any superset, e.g. the whole scene)
        if(si) CU(cudaMemcpyAsync(si, ids, idb, cudaMemcpyHostToDevice, e->stream));
        CU(cudaMemsetAsync(&e->st->n_own, 0, sizeof(int), e->stream));
        LAUNCH(e, k_upload_slab, div_up(n, 256), 256, n, sp, sv, si, dtype == BBX_F64, e->grid, e->cap, e->st, e->pos[e->cur], e->vel[e->cur], e->pid[e->cur]);
        CU(cudaGetLastError());
        rc = read_state(e); if(rc) return rc;
        if(e->st_host->error == BBX_ERR_CAPACITY) return set_error(BBX_ERR_CAPACITY, "slab holds more than max_particles = %d particles", e->cap);
        e->n = e->st_host->n_own; e->n_hint = e->n; e->hint_count = 0; e->n_launch = bound_of(e, e->n);
        return BBX_OK;
    }
    LAUNCH(e, k_upload, div_up(n, 256), 256, n, first, sp, sv, dtype == BBX_F64, e->pos[e->cur] + first, e->vel[e->cur] + first, e->pid[e->cur] + first);
    CU(cudaGetLastError());
    return BBX_OK;
}

static int set_particles(bbx_engine *e, int n, const void *pos, const void *vel, const int *ids, int dtype){
    if(n < 0 || (n > 0 && (!pos || !vel))) return set_error(BBX_ERR_INVALID, "bad particle arguments");
    if(IS_SLAB(e) && !e->comm) return set_error(BBX_ERR_INVALID, "slab engine without a communicator: call bbx_comm_init / bbx_comm_init_local first");
    if(!IS_SLAB(e) && n > e->cap) return set_error(BBX_ERR_CAPACITY, "%d particles exceed max_particles %d", n, e->cap);
    e->n = IS_SLAB(e) ? 0 : n;
    e->n_glo = e->n_ghi = e->n_first = e->n_last = 0;
    e->have_chains = 0; e->force_full = 1;
    // a fresh particle set starts with a clean slate: the sticky device error of an earlier set is dropped
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaMemsetAsync(&e->st->error, 0, sizeof(int), e->stream));
    CU(cudaMemsetAsync(&e->st->n_glo, 0, 4 * sizeof(int), e->stream)); // n_glo, n_ghi, peer_n[2]: no ghosts yet
    *e->err_probe = 0; e->st_host->error = 0; e->deferred_error = 0; e->counts_stale = 0;
    if(n == 0 && !IS_SLAB(e)) return BBX_OK;
    int rc;
    if(n > 0){ rc = upload_particles(e, 0, n, pos, vel, ids, dtype); if(rc) return rc; }
    // PciSphSolver3::Setup: initial DistributeByParticle (ascending id) so that the first sub-step can
    // run the incremental update exactly like the reference (frame_index = 1)
    rc = grid_update(e); if(rc) return rc;
    e->force_full = 0;
    return BBX_OK;
}
int bbx_set_particles(bbx_engine *e, int n, const void *pos, const void *vel, int dtype){
    CHECK_ENGINE(e);
    return set_particles(e, n, pos, vel, nullptr, dtype);
}
int bbx_set_particles_ids(bbx_engine *e, int n, const void *pos, const void *vel, const int *ids, int dtype){
    CHECK_ENGINE(e);
    if(!IS_SLAB(e) && ids) return set_error(BBX_ERR_INVALID, "explicit particle ids are only meaningful for slab engines");
    return set_particles(e, n, pos, vel, ids, dtype);
}

// Re-sort after an append: old particles keep their recorded cell and their order inside it, the appended ones
// follow in id order.  Counting sort by cell with the full-rebuild kernels, ordered by OLD SLOT (old slots are in
// chain order, the appended particles sit behind them in id order).  Not a grid epoch: flags and parity stay.
static int slab_refresh_ghosts(bbx_engine *e, int nxt, unsigned posted_seq = 0);
// split / id0: see k_full_sort_cells (slab engines: the appended particles are ordered by id, not by slot)
static int halo_settle_all(bbx_engine *e);
static int append_update(bbx_engine *e, int n_old, int k, int split = -1, int id0 = 0){
    DevGrid &g = e->grid;
    { int rc = halo_settle_all(e); if(rc) return rc; }
    const int cur = e->cur, nxt = cur ^ 1, n_all = n_old + k, par = e->epoch & 1;
    const int own_cells = g.c_own1 - g.c_own0;
    LAUNCH(e, k_append_hash, div_up(std::max(n_all, e->scan_tiles), 256), 256, n_old, k, e->pos[cur], e->cell[cur], e->newcell, e->count, g, e->st, e->scan_status, e->scan_tiles);
    LAUNCH(e, k_scan_cells, e->scan_tiles, 256, e->count + g.c_own0, own_cells, g.c_own0, e->scan_status, e->st, e->cell_start[nxt] + g.c_own0, e->occ_cells);
    LAUNCH(e, k_full_scatter, div_up(std::max(n_all, 1), 256), 256, n_all, 0, g, e->st, par, 1, e->newcell, e->cell_start[nxt], e->count, e->perm);
    LAUNCH(e, k_full_sort_cells, div_up(g.total, 256), 256, g, e->st, par, 1, e->cell_start[nxt], split >= 0 ? e->pid[cur] : (const int *)nullptr, e->perm, e->count, split, id0);
    LAUNCH(e, k_full_gather, div_up(std::max(n_all, 1), 256), 256, e->st, par, 1, e->perm, e->newcell, e->pos[cur], e->vel[cur], e->pid[cur],
           e->pos[nxt], e->vel[nxt], e->pid[nxt], e->cell[nxt], e->rec);
    CU(cudaGetLastError());
    if(IS_SLAB(e)){ int rc = slab_refresh_ghosts(e, nxt); if(rc) return rc; } // the neighbours' ghost copies of my boundary planes changed too
    e->cur = nxt; e->n = n_all; e->have_chains = 1;
    return BBX_OK;
}

int bbx_append_particles(bbx_engine *e, int n, const void *pos, const void *vel, int dtype){
    CHECK_ENGINE(e);
    if(n <= 0) return BBX_OK;
    if(IS_SLAB(e)) return set_error(BBX_ERR_INVALID, "bbx_append_particles is not available on slab engines");
    if(e->n + n > e->cap) return set_error(BBX_ERR_CAPACITY, "%d + %d particles exceed max_particles %d", e->n, n, e->cap);
    // New particles take ids n_old.. and go to the tail of their cell's chain: equivalent to a
    // stable merge; realised as "old chains first, then new ids ascending" via the incremental fill
    // followed by a tail insert.  Implemented with the full-rebuild machinery when no chains exist.
    if(!e->have_chains || e->n == 0){
        int first = e->n;
        int rc = upload_particles(e, first, n, pos, vel, nullptr, dtype); if(rc) return rc;
        e->n += n; e->force_full = 1;
        rc = grid_update(e); if(rc) return rc;
        e->force_full = 0;
        return BBX_OK;
    }
    // after stepping: ContinuousParticleSetBuilder3::Commit (grid.h:1424-1441) -- the chains of the existing
    // particles stay as they are, the new ids go to the tail of the chain of the cell they hash to
    const int n_old = e->n;
    int rc = upload_particles(e, n_old, n, pos, vel, nullptr, dtype); if(rc) return rc;
    return append_update(e, n_old, n);
}

// bbx_append_particles with explicit global ids.  Slab engines (collective over the group): every rank passes the appended
// particles (all of them, or any superset of its share) with their global ids -- which must be larger than every id
// already in the run, as the ids of ContinuousParticleSetBuilder3::AddParticle are -- and keeps those of its owned planes
// at the tail of their cells' chains, in id order; the boundary planes are exchanged again.
int bbx_append_particles_ids(bbx_engine *e, int n, const void *pos, const void *vel, const int *ids, int dtype){
    CHECK_ENGINE(e);
    if(!IS_SLAB(e)){
        if(ids) for(int k = 0; k < n; k++) if(ids[k] != e->n + k) return set_error(BBX_ERR_INVALID, "single-domain engines number appended particles themselves: ids must continue from %d", e->n);
        return bbx_append_particles(e, n, pos, vel, dtype);
    }
    if(n < 0 || (n > 0 && (!pos || !vel || !ids))) return set_error(BBX_ERR_INVALID, "bad particle arguments (slab engines need ids)");
    if(!e->have_chains) return set_error(BBX_ERR_INVALID, "append on a slab engine needs a particle set first (bbx_set_particles_ids)");
    int rc = sync_counts(e); if(rc) return rc;
    const int n_old = e->n;
    int id0 = 0x7fffffff;
    for(int k = 0; k < n; k++) id0 = std::min(id0, ids[k]);
    if(n == 0) id0 = 0;
    if(n > 0){
        if(dtype != BBX_F32 && dtype != BBX_F64) return set_error(BBX_ERR_INVALID, "dtype must be BBX_F32 or BBX_F64");
        const size_t bytes = (dtype == BBX_F64 ? 8 : 4) * 3 * (size_t)n, idb = sizeof(int) * (size_t)n;
        rc = ensure_stage(e, 2 * bytes + idb); if(rc) return rc;
        char *sp = (char *)e->stage, *sv = sp + bytes; int *si = (int *)(sv + bytes);
        CU(cudaMemcpyAsync(sp, pos, bytes, cudaMemcpyHostToDevice, e->stream));
        CU(cudaMemcpyAsync(sv, vel, bytes, cudaMemcpyHostToDevice, e->stream));
        CU(cudaMemcpyAsync(si, ids, idb, cudaMemcpyHostToDevice, e->stream));
        // the cursor st->n_own stands at the owned count: the kept particles go behind the owned slots
        LAUNCH(e, k_upload_slab, div_up(n, 256), 256, n, sp, sv, si, dtype == BBX_F64, e->grid, e->cap, e->st, e->pos[e->cur], e->vel[e->cur], e->pid[e->cur]);
        CU(cudaGetLastError());
    }
    rc = read_state(e); if(rc) return rc;
    if(e->st_host->error == BBX_ERR_CAPACITY) return set_error(BBX_ERR_CAPACITY, "slab would hold more than max_particles = %d particles", e->cap);
    const int kept = e->st_host->n_own - n_old;
    e->n_hint = n_old + kept; e->hint_count = 0; e->n_launch = std::max(e->n_launch, bound_of(e, n_old + kept)); // (k_slab_plan checks the count against it)
    return append_update(e, n_old, kept, n_old, id0);
}

int bbx_particle_count(bbx_engine *e, int *n){
    if(!e || !n) return set_error(BBX_ERR_INVALID, "null");
    if(e->counts_stale){ CU(cudaSetDevice(e->device)); int rc = sync_counts(e); if(rc) return rc; }
    *n = e->n; return BBX_OK;
}

int bbx_overwrite_state(bbx_engine *e, const void *pos, const void *vel, int dtype){
    CHECK_ENGINE(e);
    if(IS_SLAB(e)) return set_error(BBX_ERR_INVALID, "bbx_overwrite_state is not available on slab engines");
    if(dtype != BBX_F32 && dtype != BBX_F64) return set_error(BBX_ERR_INVALID, "dtype must be BBX_F32 or BBX_F64");
    if(e->n == 0) return BBX_OK;
    if(!pos || !vel) return set_error(BBX_ERR_INVALID, "null positions / velocities (bbx_particle_count rows of 3 values each are read)");
    size_t esz = dtype == BBX_F64 ? 8 : 4; size_t bytes = esz * 3 * (size_t)e->n;
    int rc = ensure_stage(e, 2 * bytes); if(rc) return rc;
    char *sp = (char *)e->stage, *sv = sp + bytes;
    CU(cudaMemcpyAsync(sp, pos, bytes, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(sv, vel, bytes, cudaMemcpyHostToDevice, e->stream));
    LAUNCH(e, k_overwrite, div_up(e->n, 256), 256, e->n, e->pid[e->cur], sp, sv, dtype == BBX_F64, e->pos[e->cur], e->vel[e->cur]);
    CU(cudaGetLastError());
    return BBX_OK;
}

static int exchange2(bbx_engine *e, float4 *a, float4 *b);
int bbx_overwrite_owned(bbx_engine *e, const void *pos, const void *vel, int dtype){
    CHECK_ENGINE(e);
    { int rc_ = sync_counts(e); if(rc_) return rc_; }
    if(dtype != BBX_F32 && dtype != BBX_F64) return set_error(BBX_ERR_INVALID, "dtype must be BBX_F32 or BBX_F64");
    if(e->n > 0){
        if(!pos || !vel) return set_error(BBX_ERR_INVALID, "null");
        size_t esz = dtype == BBX_F64 ? 8 : 4; size_t bytes = esz * 3 * (size_t)e->n;
        int rc = ensure_stage(e, 2 * bytes); if(rc) return rc;
        char *sp = (char *)e->stage, *sv = sp + bytes;
        CU(cudaMemcpyAsync(sp, pos, bytes, cudaMemcpyHostToDevice, e->stream));
        CU(cudaMemcpyAsync(sv, vel, bytes, cudaMemcpyHostToDevice, e->stream));
        LAUNCH(e, k_overwrite, div_up(e->n, 256), 256, e->n, (const int *)nullptr, sp, sv, dtype == BBX_F64, e->pos[e->cur], e->vel[e->cur]);
        CU(cudaGetLastError());
    }
    // the neighbours run their next grid update on their ghost copies of my boundary planes (old order)
    return exchange2(e, e->pos[e->cur], e->vel[e->cur]);
}

// ------------------------------------------------------------------------------------ colliders
// BVH over the triangles of a mesh collider (host): median split of the centroids along the widest axis, leaves of <= 4
// triangles.  Any BVH gives the same closest distance (the traversal only skips boxes that cannot beat the best so far).
struct BvhTri { double lo[3], hi[3], ctr[3]; int id; };
static int bvh_build(std::vector<DevBvhNode> &nodes, std::vector<BvhTri> &tris, int begin, int end){
    const int me = (int)nodes.size();
    nodes.push_back(DevBvhNode());
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, clo[3] = {1e300, 1e300, 1e300}, chi[3] = {-1e300, -1e300, -1e300};
    for(int t = begin; t < end; t++) for(int k = 0; k < 3; k++){
        lo[k] = std::min(lo[k], tris[t].lo[k]); hi[k] = std::max(hi[k], tris[t].hi[k]);
        clo[k] = std::min(clo[k], tris[t].ctr[k]); chi[k] = std::max(chi[k], tris[t].ctr[k]);
    }
    DevBvhNode n; memset(&n, 0, sizeof(n));
    for(int k = 0; k < 3; k++){ n.lo[k] = lo[k]; n.hi[k] = hi[k]; }
    n.left = n.right = -1; n.first = begin; n.count = 0;
    if(end - begin <= 4){ n.count = end - begin; nodes[me] = n; return me; }
    int axis = 0; for(int k = 1; k < 3; k++) if(chi[k] - clo[k] > chi[axis] - clo[axis]) axis = k;
    const int mid = (begin + end) / 2;
    std::nth_element(tris.begin() + begin, tris.begin() + mid, tris.begin() + end, [axis](const BvhTri &a, const BvhTri &b){ return a.ctr[axis] < b.ctr[axis]; });
    n.left = bvh_build(nodes, tris, begin, mid);
    n.right = bvh_build(nodes, tris, mid, end);
    nodes[me] = n;
    return me;
}
static int fill_mesh(bbx_engine *e, DevCollider &d, const bbx_collider &c){
    if(c.mesh_vertices <= 0 || c.mesh_triangles <= 0 || !c.mesh_points || !c.mesh_indices) return set_error(BBX_ERR_INVALID, "mesh collider without triangles");
    const int nv = c.mesh_vertices, nt = c.mesh_triangles;
    std::vector<BvhTri> tris((size_t)nt);
    for(int t = 0; t < nt; t++){
        BvhTri &b = tris[t]; b.id = t;
        for(int k = 0; k < 3; k++){ b.lo[k] = 1e300; b.hi[k] = -1e300; b.ctr[k] = 0; }
        for(int v = 0; v < 3; v++){
            const int ix = c.mesh_indices[3 * (size_t)t + v];
            if(ix < 0 || ix >= nv) return set_error(BBX_ERR_INVALID, "mesh index %d out of range (%d vertices)", ix, nv);
            for(int k = 0; k < 3; k++){ const double x = c.mesh_points[3 * (size_t)ix + k]; b.lo[k] = std::min(b.lo[k], x); b.hi[k] = std::max(b.hi[k], x); b.ctr[k] += x / 3.0; }
        }
    }
    std::vector<DevBvhNode> nodes; nodes.reserve((size_t)nt);
    bvh_build(nodes, tris, 0, nt);
    std::vector<int> order(3 * (size_t)nt);
    for(int t = 0; t < nt; t++) for(int v = 0; v < 3; v++) order[3 * (size_t)t + v] = c.mesh_indices[3 * (size_t)tris[t].id + v];
    double *dp = nullptr; int *di = nullptr; DevBvhNode *dn = nullptr;
    CU(cudaMalloc((void **)&dp, sizeof(double) * 3 * (size_t)nv)); e->mesh_allocs.push_back(dp);
    CU(cudaMalloc((void **)&di, sizeof(int) * 3 * (size_t)nt)); e->mesh_allocs.push_back(di);
    CU(cudaMalloc((void **)&dn, sizeof(DevBvhNode) * nodes.size())); e->mesh_allocs.push_back(dn);
    CU(cudaMemcpy(dp, c.mesh_points, sizeof(double) * 3 * (size_t)nv, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(di, order.data(), sizeof(int) * 3 * (size_t)nt, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dn, nodes.data(), sizeof(DevBvhNode) * nodes.size(), cudaMemcpyHostToDevice));
    d.mesh_points = dp; d.mesh_tris = di; d.bvh = dn; d.n_tris = nt; d.n_nodes = (int)nodes.size();
    for(int k = 0; k < 3; k++){ d.mesh_lo[k] = nodes[0].lo[k]; d.mesh_hi[k] = nodes[0].hi[k]; } // MeshGetBounds = the BVH root
    return BBX_OK;
}

static int fill_collider(bbx_engine *e, DevCollider &d, const bbx_collider &c){
    if(c.type < BBX_COLLIDER_BOX || c.type > BBX_COLLIDER_MESH) return set_error(BBX_ERR_INVALID, "unknown collider type %d", c.type);
    d.type = c.type; d.reverse = c.reverse_orientation ? 1 : 0; d.active = c.active ? 1 : 0;
    memcpy(d.o2w, c.object_to_world, sizeof(d.o2w)); memcpy(d.w2o, c.world_to_object, sizeof(d.w2o));
    memcpy(d.size, c.size, sizeof(d.size)); d.radius = c.radius; d.friction = c.friction;
    memcpy(d.linvel, c.linear_velocity, sizeof(d.linvel)); memcpy(d.angvel, c.angular_velocity, sizeof(d.angvel));
    d.sdf_field = nullptr; d.sdf_field32 = nullptr; d.lipschitz = 0.0;
    d.mesh_points = nullptr; d.mesh_tris = nullptr; d.bvh = nullptr; d.n_tris = d.n_nodes = 0;
    if(c.type == BBX_COLLIDER_MESH){ int rc = fill_mesh(e, d, c); if(rc) return rc; }
    if(c.type == BBX_COLLIDER_SDF || c.type == BBX_COLLIDER_MESH){
        if(!c.sdf_field) return set_error(BBX_ERR_INVALID, c.type == BBX_COLLIDER_MESH ? "mesh collider without its SDF grid" : "SDF collider without field");
        size_t total = (size_t)c.sdf_resolution[0] * c.sdf_resolution[1] * c.sdf_resolution[2];
        if(total == 0) return set_error(BBX_ERR_INVALID, "SDF collider with empty grid");
        double *dev = nullptr;
        CU(cudaMalloc((void **)&dev, total * sizeof(double)));
        CU(cudaMemcpy(dev, c.sdf_field, total * sizeof(double), cudaMemcpyHostToDevice));
        e->sdf_fields.push_back(dev);
        d.sdf_field = dev;
        for(int k = 0; k < 3; k++){ d.sdf_res[k] = c.sdf_resolution[k]; d.sdf_spacing[k] = c.sdf_spacing[k]; d.sdf_origin[k] = c.sdf_origin[k]; }
        // FP32 shadow + Lipschitz bound of the trilinear field for the conservative pre-check (bbx_cull)
        std::vector<float> f32(total);
        double gmax[3] = {0, 0, 0};
        const int rx = c.sdf_resolution[0], ry = c.sdf_resolution[1], rz = c.sdf_resolution[2];
        for(int z = 0; z < rz; z++) for(int y = 0; y < ry; y++) for(int x = 0; x < rx; x++){
            size_t id = (size_t)x + (size_t)y * rx + (size_t)z * rx * ry;
            double v = c.sdf_field[id];
            f32[id] = (float)v;
            if(x + 1 < rx) gmax[0] = std::max(gmax[0], fabs(c.sdf_field[id + 1] - v) / c.sdf_spacing[0]);
            if(y + 1 < ry) gmax[1] = std::max(gmax[1], fabs(c.sdf_field[id + rx] - v) / c.sdf_spacing[1]);
            if(z + 1 < rz) gmax[2] = std::max(gmax[2], fabs(c.sdf_field[id + (size_t)rx * ry] - v) / c.sdf_spacing[2]);
        }
        float *dev32 = nullptr;
        CU(cudaMalloc((void **)&dev32, total * sizeof(float)));
        CU(cudaMemcpy(dev32, f32.data(), total * sizeof(float), cudaMemcpyHostToDevice));
        e->sdf_fields32.push_back(dev32);
        d.lipschitz = sqrt(gmax[0] * gmax[0] + gmax[1] * gmax[1] + gmax[2] * gmax[2]);
        d.sdf_field32 = dev32;
    }
    return BBX_OK;
}
// FP32 shadow of the collider set (+ domain faces) for the conservative pre-check of the sweeps
static int push_cull(bbx_engine *e){
    DevCullSet &q = e->cull_host;
    memset(&q, 0, sizeof(q));
    const DevColliderSet &cs = e->colliders_host;
    q.count = cs.count;
    double scale = 1.0;
    for(int k = 0; k < 3; k++) scale = std::max(scale, std::max(fabs(e->grid.min[k]), fabs(e->grid.max[k])));
    for(int i = 0; i < cs.count; i++){
        const DevCollider &d = cs.c[i];
        DevCullCollider &c = q.c[i];
        c.type = d.type; c.reverse = d.reverse; c.active = d.active;
        bool ident = true;
        for(int r = 0; r < 3; r++) for(int k = 0; k < 4; k++){
            c.w2o[r * 4 + k] = (float)d.w2o[r * 4 + k];
            if(d.w2o[r * 4 + k] != (r == k ? 1.0 : 0.0)) ident = false;
            scale = std::max(scale, fabs(d.w2o[r * 4 + k]) * (k == 3 ? 1.0 : 0.0));
        }
        // a projective last row would need the divide of Transform::Point: never cull such a collider
        bool affine = d.w2o[12] == 0 && d.w2o[13] == 0 && d.w2o[14] == 0 && d.w2o[15] == 1;
        c.identity = ident ? 1 : 0;
        for(int k = 0; k < 3; k++){ c.half[k] = (float)(d.size[k] / 2.0); scale = std::max(scale, d.size[k]); }
        c.radius = (float)d.radius; scale = std::max(scale, d.radius);
        c.lipschitz = (float)(d.lipschitz * (1.0 + 1e-5) + 1e-12);
        for(int k = 0; k < 3; k++){
            c.sdf_res[k] = d.sdf_res[k];
            c.sdf_inv_spacing[k] = d.sdf_spacing[k] > 0 ? (float)(1.0 / d.sdf_spacing[k]) : 0.f;
            c.sdf_origin[k] = (float)d.sdf_origin[k];
        }
        c.sdf_field32 = d.sdf_field32;
        if(!affine) c.type = -1; // never cleared by the pre-check
    }
    // FP32 evaluation error of the signed distances: a few ulp of the largest coordinate involved
    q.margin = (float)(1.0e-5 * scale + 1.0e-3 * e->cfg.spacing);
    for(int k = 0; k < 3; k++){ q.dom_lo[k] = (float)(e->grid.min[k] + q.margin); q.dom_hi[k] = (float)(e->grid.max[k] - q.margin); }
    CU(cudaMemcpyAsync(e->cull, &q, sizeof(DevCullSet), cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return BBX_OK;
}
static int push_colliders(bbx_engine *e){
    CU(cudaMemcpyAsync(e->colliders, &e->colliders_host, sizeof(DevColliderSet), cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream)); // colliders_host may change right after this call
    return push_cull(e);
}
int bbx_set_colliders(bbx_engine *e, int n, const bbx_collider *colliders){
    CHECK_ENGINE(e);
    if(n < 0 || n > BBX_MAX_COLLIDERS || (n > 0 && !colliders)) return set_error(BBX_ERR_INVALID, "0..%d colliders supported", BBX_MAX_COLLIDERS);
    CU(cudaStreamSynchronize(e->stream));
    for(double *f : e->sdf_fields) cudaFree(f);
    e->sdf_fields.clear();
    for(float *f : e->sdf_fields32) cudaFree(f);
    e->sdf_fields32.clear();
    for(void *m : e->mesh_allocs) cudaFree(m);
    e->mesh_allocs.clear();
    memset(&e->colliders_host, 0, sizeof(DevColliderSet));
    for(int i = 0; i < n; i++){ int rc = fill_collider(e, e->colliders_host.c[i], colliders[i]); if(rc) return rc; }
    e->colliders_host.count = n;
    return push_colliders(e);
}
int bbx_update_collider(bbx_engine *e, int index, const bbx_collider *c){
    CHECK_ENGINE(e);
    if(!c || index < 0 || index >= e->colliders_host.count) return set_error(BBX_ERR_INVALID, "collider index out of range");
    DevCollider &d = e->colliders_host.c[index];
    if(c->type != d.type) return set_error(BBX_ERR_INVALID, "bbx_update_collider cannot change the collider type");
    if(d.type == BBX_COLLIDER_MESH && (memcmp(d.o2w, c->object_to_world, sizeof(d.o2w)) != 0))
        return set_error(BBX_ERR_INVALID, "a mesh collider is stored in world space with its baked SDF: it cannot be moved, set the colliders again");
    const double *keep = d.sdf_field; // the baked field (and its FP32 shadow, Lipschitz bound) stay
    d.reverse = c->reverse_orientation ? 1 : 0; d.active = c->active ? 1 : 0;
    memcpy(d.o2w, c->object_to_world, sizeof(d.o2w)); memcpy(d.w2o, c->world_to_object, sizeof(d.w2o));
    memcpy(d.size, c->size, sizeof(d.size)); d.radius = c->radius; d.friction = c->friction;
    memcpy(d.linvel, c->linear_velocity, sizeof(d.linvel)); memcpy(d.angvel, c->angular_velocity, sizeof(d.angvel));
    d.sdf_field = keep;
    return push_colliders(e);
}
int bbx_collider_distance(bbx_engine *e, int index, int n, const double *points, double *out){
    CHECK_ENGINE(e);
    if(index < 0 || index >= e->colliders_host.count) return set_error(BBX_ERR_INVALID, "collider index out of range");
    if(n <= 0) return BBX_OK;
    if(!points || !out) return set_error(BBX_ERR_INVALID, "null");
    const size_t bp = sizeof(double) * 3 * (size_t)n, bo = sizeof(double) * (size_t)n;
    int rc = ensure_stage(e, bp + bo); if(rc) return rc;
    double *dp = (double *)e->stage, *dout = dp + 3 * (size_t)n;
    CU(cudaMemcpyAsync(dp, points, bp, cudaMemcpyHostToDevice, e->stream));
    LAUNCH(e, k_collider_distance, div_up(n, 128), 128, n, e->colliders, index, dp, dout);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, dout, bo, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return BBX_OK;
}
int bbx_set_collider_active(bbx_engine *e, int index, int active){
    CHECK_ENGINE(e);
    if(index < 0 || index >= e->colliders_host.count) return set_error(BBX_ERR_INVALID, "collider index out of range");
    e->colliders_host.c[index].active = active ? 1 : 0;
    return push_colliders(e);
}

// --------------------------------------------------------------------------------------- timing
static void tick(bbx_engine *e, int phase){
    if(!e->timing) return;
    if(e->ev_used == e->ev.size()){ cudaEvent_t ev; cudaEventCreate(&ev); e->ev.push_back(ev); e->ev_phase.push_back(0); }
    e->ev_phase[e->ev_used] = phase;
    cudaEventRecord(e->ev[e->ev_used++], e->stream);
}
// events are recorded as (phase start) ... (T_COUNT = end of step); harvest after a sync
static void harvest(bbx_engine *e){
    if(!e->timing || e->ev_used < 2) { e->ev_used = 0; return; }
    cudaStreamSynchronize(e->stream);
    float step_ms = 0.f, grid_ms = 0.f;
    for(size_t k = 0; k + 1 < e->ev_used; k++){
        int ph = e->ev_phase[k];
        float ms = 0.f; cudaEventElapsedTime(&ms, e->ev[k], e->ev[k + 1]);
        if(ph >= T_COUNT){ step_ms = 0.f; grid_ms = 0.f; e->phase_ms[T_COUNT] += ms; e->phase_launches[T_COUNT]++; continue; } // gap between sub-steps
        e->phase_ms[ph] += ms; e->phase_launches[ph]++;
        step_ms += ms; if(ph == T_GRID) grid_ms += ms;
        e->last_ms_step = step_ms; e->last_ms_grid = grid_ms;
    }
    e->ev_used = 0;
}

// ------------------------------------------------------------------------------------- stepping
static void make_params(bbx_engine *e, double dt, StepParams &P){
    const bbx_config &c = e->cfg;
    double h = e->h, pi = 3.14159265358979323846;
    P.n = launch_n(e); P.n_owned = P.n; P.dyn = IS_SLAB(e) ? e->st : nullptr;
    P.h = (float)h; P.h2 = (float)(h * h); P.inv_h = (float)(1.0 / h); P.inv_h2 = (float)(1.0 / (h * h));
    P.h_d = h; P.h2_d = h * h;
    P.thr2 = (float)(h * h - 1e-8);
    P.band = (float)(h * h * 1.0e-6 + 4e-8); // >> FP32 rounding of d^2 (~1e-7 relative) and of thr2
    P.thr_lo = P.thr2 - P.band; P.thr_hi = P.thr2 + P.band;
    // list build: x = 1 - d^2 / h^2 evaluated in the cell frame (absolute error of a few 1e-7, bbx_lists.cuh)
    { double bx = 4.0e-6 + 4e-8 / (h * h); P.xacc = (float)(1e-8 / (h * h) - bx); P.xband = (float)(1e-8 / (h * h) + bx); }
    P.par = (e->epoch + 1) & 1; // parity of the grid epoch this sub-step runs under (set right after its grid update)
    P.part = 0;
    P.mass = (float)e->mass; P.mass2 = (float)(e->mass * e->mass); P.inv_mass = (float)(1.0 / e->mass);
    P.rho0 = (float)c.target_density;
    P.w_std_c = (float)(315.0 / (64.0 * pi * h * h * h));
    P.d2w_spiky_c = (float)(90.0 / (pi * h * h * h * h * h));
    P.dw_spiky_c = (float)(45.0 / (pi * h * h * h * h));
    P.w_spiky_c = (float)(15.0 / (pi * h * h * h));
    P.viscosity = (float)c.viscosity; P.drag = (float)c.drag;
    P.gx = (float)c.gravity[0]; P.gy = (float)c.gravity[1]; P.gz = (float)c.gravity[2];
    P.dt = (float)dt;
    P.delta = (float)bbxh_delta(e->mass_over_rho0_sq, e->delta_denom, dt);
    P.neg_pressure_scale = (float)c.negative_pressure_scale;
    P.eos_scale = (float)(c.target_density * c.sound_speed * c.sound_speed / c.eos_exponent);
    P.eos_exponent = (float)c.eos_exponent;
    P.radius = (float)c.spacing;
    P.restitution = (float)c.restitution;
    double minlen = std::min(e->grid.len[0], std::min(e->grid.len[1], e->grid.len[2]));
    P.min_cell_len09 = (float)(0.9 * minlen);
    double pf = dt * c.pseudo_viscosity; P.pseudo_factor = (float)(pf < 0 ? 0 : (pf > 1 ? 1 : pf));
}

// ---- slab engines: exchange the boundary planes of slot-indexed arrays with the two neighbours.
// Send: my first owned plane [0, n_first) to the lower neighbour, my last owned plane [n - n_last, n) to
// the upper one.  Receive: the lower ghost plane [-n_glo, 0) and the upper one [n, n + n_ghi).  The ranges
// are contiguous because slots are sorted by cell id and z is the slowest cell index -- no pack kernels.
static int exchange_planes(bbx_engine *e, void *const *arr, const size_t *esz, int narr){
    if(!IS_SLAB(e)) return BBX_OK;
    { int rc = halo_settle_all(e); if(rc) return rc; }
    { int rc = sync_counts(e); if(rc) return rc; } // (send / recv sizes are host values)
    BbxSeg slo[BBX_MAX_SEGS], rlo[BBX_MAX_SEGS], shi[BBX_MAX_SEGS], rhi[BBX_MAX_SEGS];
    for(int k = 0; k < narr; k++){
        char *base = (char *)arr[k]; size_t z = esz[k];
        slo[k].ptr = base; slo[k].bytes = z * (size_t)e->n_first;
        rlo[k].ptr = base - z * (size_t)e->n_glo; rlo[k].bytes = z * (size_t)e->n_glo;
        shi[k].ptr = base + z * (size_t)(e->n - e->n_last); shi[k].bytes = z * (size_t)e->n_last;
        rhi[k].ptr = base + z * (size_t)e->n; rhi[k].bytes = z * (size_t)e->n_ghi;
    }
    COMM(e->comm->exchange(e->stream, slo, rlo, narr, shi, rhi, narr));
    return BBX_OK;
}
static int exchange1(bbx_engine *e, float4 *a){ void *arr[1] = {a}; size_t z[1] = {sizeof(float4)}; return exchange_planes(e, arr, z, 1); }
static int exchange2(bbx_engine *e, float4 *a, float4 *b){ void *arr[2] = {a, b}; size_t z[2] = {sizeof(float4), sizeof(float4)}; return exchange_planes(e, arr, z, 2); }

// ---- halo push: where the boundary planes of array(s) `which` go in the neighbours' memory
enum { HALO_DENSITY = 0, HALO_PREDICT, HALO_PRESSURE, HALO_INTEGRATE, HALO_COUNTS, HALO_PLANES };
static HaloDst halo_none(){ HaloDst h; memset(&h, 0, sizeof(h)); h.hi_begin = 0x7fffffff; return h; }
static HaloDst halo_dst(bbx_engine *e, float4 *lo0, float4 *hi0, float4 *lo1 = nullptr, float4 *hi1 = nullptr){
    HaloDst h = halo_none();
    if(!e->p2p) return h;
    h.lo[0] = lo0; h.hi[0] = hi0; h.lo[1] = lo1; h.hi[1] = hi1;
    h.has_lo = e->has_lo; h.has_hi = e->has_hi; // the plane sizes are resolved on the device (bbx_halo_resolve)
    return h;
}
// end of a phase whose kernels pushed their boundary planes: raise my flag at both neighbours, wait for theirs
static unsigned *halo_flag_at(bbx_engine *e, int side, int phase){ // the flag I raise at neighbour `side` (I am its opposite side)
    if(side == 0 ? !e->has_lo : !e->has_hi) return nullptr;
    return e->peer[side].flags + (side == 0 ? 1 : 0) * BBX_HALO_PHASES + phase;
}
static int halo_wait(bbx_engine *e, int phase, unsigned seq, cudaStream_t strm = nullptr){
    LAUNCH_ON(e, strm ? strm : e->stream, k_halo_wait, 1, 32, 0, e->has_lo ? e->halo_flags + 0 * BBX_HALO_PHASES + phase : (const unsigned *)nullptr,
           e->has_hi ? e->halo_flags + 1 * BBX_HALO_PHASES + phase : (const unsigned *)nullptr, seq, &e->st->error);
    CU(cudaGetLastError());
    return BBX_OK;
}
static int halo_sync(bbx_engine *e, int phase){
    const unsigned seq = ++e->halo_seq[phase];
    LAUNCH(e, k_halo_signal, 1, 32, halo_flag_at(e, 0, phase), halo_flag_at(e, 1, phase), seq);
    return halo_wait(e, phase, seq);
}
// Overlapped form (slab engines with the halo push, reference-compat PCISPH step).  A phase runs as two launches: its
// BOUNDARY blocks (every block holding a slot of the first / last owned plane) on `bstream`, its interior blocks on
// `stream`, side by side.  The boundary blocks store their results straight into the neighbours and raise the phase flag;
// the matching wait sits in front of the boundary blocks of the NEXT phase only -- interior blocks never touch a ghost
// slot.  So a rank that runs ahead of its neighbour by less than an interior sweep never idles, instead of paying the
// skew phase by phase.  Dependencies inside the rank: boundary(p) after interior(p-1) [event], interior(p) after
// boundary(p-1) [event]; the two streams order the rest.
static void halo_signal(bbx_engine *e, int phase, cudaStream_t strm){
    const unsigned seq = ++e->halo_seq[phase];
    LAUNCH_ON(e, strm, k_halo_signal, 1, 32, 0, halo_flag_at(e, 0, phase), halo_flag_at(e, 1, phase), seq);
    e->halo_due[phase] = seq;
}
static int halo_settle(bbx_engine *e, int phase, cudaStream_t strm = nullptr){
    if(!e->halo_due[phase]) return BBX_OK;
    const unsigned seq = e->halo_due[phase];
    e->halo_due[phase] = 0;
    return halo_wait(e, phase, seq, strm);
}
static int halo_settle_all(bbx_engine *e){
    for(int ph = HALO_DENSITY; ph <= HALO_INTEGRATE; ph++){ int rc = halo_settle(e, ph); if(rc) return rc; }
    return BBX_OK;
}
// start of overlapped phase `ph` (HALO_DENSITY .. HALO_INTEGRATE): order the two streams against the previous phase and
// put the wait for the neighbours' previous halo in front of the boundary blocks
static int ov_begin(bbx_engine *e, int ph){
    if(ph == HALO_DENSITY){
        CU(cudaEventRecord(e->ev_grid, e->stream));                       // the grid update (ghost planes included) is on `stream`
        CU(cudaStreamWaitEvent(e->bstream, e->ev_grid, 0));
        return BBX_OK;
    }
    CU(cudaStreamWaitEvent(e->bstream, e->ev_i[ph - 1], 0));
    CU(cudaStreamWaitEvent(e->stream, e->ev_b[ph - 1], 0));
    return halo_settle(e, ph - 1, e->bstream);
}
static int ov_end(bbx_engine *e, int ph){
    CU(cudaEventRecord(e->ev_b[ph], e->bstream));
    CU(cudaEventRecord(e->ev_i[ph], e->stream));
    if(ph == HALO_INTEGRATE) CU(cudaStreamWaitEvent(e->stream, e->ev_b[ph], 0));   // `stream` done = sub-step done
    return BBX_OK;
}
// after a single phase run on its own (bbx_run_phase): whatever follows on `stream` sees the boundary blocks' results too
static int ov_join(bbx_engine *e){
    if(!e->overlap) return BBX_OK;
    CU(cudaEventRecord(e->ev_grid, e->bstream));
    CU(cudaStreamWaitEvent(e->stream, e->ev_grid, 0));
    return BBX_OK;
}
// blocks of T slots that can hold slots of the boundary planes (upper bound: a boundary plane fits the neighbour's ghost slots)
static int boundary_blocks(bbx_engine *e, int T){
    long long slots = 0;
    if(e->has_lo) slots += std::min<long long>(e->cap, e->peer[0].gc);
    if(e->has_hi) slots += std::min<long long>(e->cap, e->peer[1].gc);
    slots = std::min<long long>(slots, launch_n(e));
    return (int)(slots / T) + 4;
}
static inline StepParams with_part(const StepParams &P, int part){ StepParams Q = P; Q.part = part; return Q; }
// once the communicator is up: map the neighbours' arrays (BBX_P2P=0 keeps the send / recv path)
static int setup_peers(bbx_engine *e){
    e->p2p = 0; e->peers_ready = 1;
    if(!IS_SLAB(e) || !e->comm) return BBX_OK;
    const char *env = getenv("BBX_P2P");
    const int want = (env && env[0] == '0') ? 0 : 1;
    BbxPeerArrays mine, lo, hi; int ok = 0;
    for(int k = 0; k < BBX_PEER_NPTR; k++) mine.p[k] = e->raw_shared[k];
    mine.gc = e->gcap;
    COMM(e->comm->share_arrays(e->stream, mine, want, &lo, &hi, &ok));
    if(!ok) return BBX_OK;
    const BbxPeerArrays *src[2] = {&lo, &hi};
    for(int sd = 0; sd < 2; sd++){
        if(sd == 0 ? !e->has_lo : !e->has_hi) continue;
        const BbxPeerArrays &a = *src[sd]; const size_t gc = (size_t)a.gc;
        bbx_engine::PeerSide &p = e->peer[sd];
        p.pos[0] = (float4 *)a.p[0] + gc; p.pos[1] = (float4 *)a.p[1] + gc; p.vel[0] = (float4 *)a.p[2] + gc; p.vel[1] = (float4 *)a.p[3] + gc;
        p.rec = (float4 *)a.p[4] + 2 * gc; p.pred = (float4 *)a.p[5] + gc; p.posq = (float4 *)a.p[6] + gc;
        p.flags = (unsigned *)a.p[7];
        p.pid[0] = (int *)a.p[8] + gc; p.pid[1] = (int *)a.p[9] + gc; p.gtab = (int *)a.p[10];
        p.gc = a.gc;
    }
    e->p2p = 1;
    {
        const char *ov = getenv("BBX_OVERLAP");
#ifdef BBX_LISTS_V7
        e->overlap = 0;
#else
        e->overlap = (ov && ov[0] == '0') ? 0 : 1;
#endif
    }
    return BBX_OK;
}

// Slab engines, after the owned slots of buffer `nxt` have been (re)ordered: the ghost planes of that buffer are replaced by
// the neighbours' freshly ordered boundary planes (collective: counts, planes, ghost part of the cell table).
// (halo push) my boundary-plane sizes and owned count into the neighbours' mailboxes + the counts flag: possible as soon as
// the scan has produced the new cell table -- the grid update posts them BEFORE its fill, so that the neighbours' counts are
// on their way while the fill runs (returns the sequence number slab_refresh_ghosts waits for)
static unsigned post_counts(bbx_engine *e, int nxt){
    DevGrid &g = e->grid;
    int *mlo = e->has_lo ? (int *)(e->peer[0].flags + 2 * BBX_HALO_PHASES) + 1 * BBX_HALO_MAIL : nullptr; // I am its UPPER side
    int *mhi = e->has_hi ? (int *)(e->peer[1].flags + 2 * BBX_HALO_PHASES) + 0 * BBX_HALO_MAIL : nullptr;
    const unsigned seq = ++e->halo_seq[HALO_COUNTS];
    LAUNCH(e, k_slab_counts, 1, 1, g, e->st, e->cell_start[nxt], e->has_lo, e->has_hi, mlo, mhi, halo_flag_at(e, 0, HALO_COUNTS), halo_flag_at(e, 1, HALO_COUNTS), seq);
    return seq;
}
static int slab_refresh_ghosts(bbx_engine *e, int nxt, unsigned posted_seq){
    DevGrid &g = e->grid;
    // migration happened implicitly: particles that crossed into my planes were found in my ghost
    // planes' old chains (in the reference's order), particles that left simply were not placed.
    // Now the ghost planes are replaced by the neighbours' freshly ordered boundary planes.
    const size_t tb = sizeof(int) * ((size_t)g.plane + 1);
    if(e->p2p){
        // sizes travel through the neighbours' mailboxes (peer memory) and stay on the device: k_slab_counts posts mine,
        // k_slab_plan reads theirs into DevState, k_push_planes takes its ranges from there -- no host round trip
        int *mail = (int *)(e->halo_flags + 2 * BBX_HALO_PHASES);
        const unsigned seq = posted_seq ? posted_seq : post_counts(e, nxt);
        int rc = halo_wait(e, HALO_COUNTS, seq); if(rc) return rc;
        LAUNCH(e, k_slab_plan, 1, 1, e->st, mail, e->has_lo, e->has_hi, (int)std::min<long long>(e->peer[0].gc, 0x7fffffff), (int)std::min<long long>(e->peer[1].gc, 0x7fffffff), e->n_launch);
        {   // the counts of this update on their way to the host (looked at BBX_HINT_LAG updates from now)
            const int slot = (int)(e->hint_count % BBX_HINT_RING);
            CU(cudaMemcpyAsync(&e->st_hint[slot], e->st, sizeof(DevState), cudaMemcpyDeviceToHost, e->stream));
            CU(cudaEventRecord(e->ev_hint[slot], e->stream));
            e->hint_count++;
        }
        // my freshly ordered boundary planes -> the neighbours' ghost slots (their table slices: lower ghost
        // plane at gtab, upper one at gtab + plane + 1), then the planes flag
        PushPtrs Q; memset(&Q, 0, sizeof(Q));
        Q.has_lo = e->has_lo; Q.has_hi = e->has_hi; Q.plane = g.plane;
        Q.tab_first = e->cell_start[nxt] + g.c_own0; Q.tab_last = e->cell_start[nxt] + g.c_own1 - g.plane;
        Q.pos = e->pos[nxt]; Q.vel = e->vel[nxt]; Q.pid = e->pid[nxt];
        if(e->has_lo){ const bbx_engine::PeerSide &p = e->peer[0]; Q.gtab_lo = p.gtab + g.plane + 1; Q.pos_lo = p.pos[nxt]; Q.vel_lo = p.vel[nxt]; Q.pid_lo = p.pid[nxt]; }
        if(e->has_hi){ const bbx_engine::PeerSide &p = e->peer[1]; Q.gtab_hi = p.gtab; Q.pos_hi = p.pos[nxt]; Q.vel_hi = p.vel[nxt]; Q.pid_hi = p.pid[nxt]; }
        LAUNCH(e, k_push_planes, e->sm_count, 256, Q, e->st);
        rc = halo_sync(e, HALO_PLANES); if(rc) return rc;
        e->counts_stale = 1;
    }else{
        LAUNCH(e, k_slab_counts, 1, 1, g, e->st, e->cell_start[nxt], e->has_lo, e->has_hi, (int *)nullptr, (int *)nullptr, (unsigned *)nullptr, (unsigned *)nullptr, 0u);
        int rc = read_state(e); if(rc) return rc;
        const int n_new = e->st_host->n_own, nf = e->st_host->n_first, nl = e->st_host->n_last;
        // boundary-plane sizes (the next exchange) and owned counts
        const int to_lo[BBX_NCOUNTS] = {nf, n_new}, to_hi[BBX_NCOUNTS] = {nl, n_new};
        int from_lo[BBX_NCOUNTS], from_hi[BBX_NCOUNTS];
        COMM(e->comm->neighbor_counts(e->stream, to_lo, to_hi, from_lo, from_hi));
        const int glo = from_lo[0], ghi = from_hi[0]; e->peer[0].n = from_lo[1]; e->peer[1].n = from_hi[1];
        if(n_new > e->cap) return set_error(BBX_ERR_CAPACITY, "slab now owns %d particles, max_particles is %d", n_new, e->cap);
        if(glo > e->gcap || ghi > e->gcap) return set_error(BBX_ERR_CAPACITY, "ghost plane of %d particles exceeds ghost_capacity %d", std::max(glo, ghi), e->gcap);
        LAUNCH(e, k_slab_plan_host, 1, 1, e->st, glo, ghi, e->peer[0].n, e->peer[1].n);
        BbxSeg slo[4] = {{e->cell_start[nxt] + g.c_own0, tb}, {e->pos[nxt], sizeof(float4) * (size_t)nf}, {e->vel[nxt], sizeof(float4) * (size_t)nf}, {e->pid[nxt], sizeof(int) * (size_t)nf}};
        BbxSeg rlo[4] = {{e->gtab, tb}, {e->pos[nxt] - glo, sizeof(float4) * (size_t)glo}, {e->vel[nxt] - glo, sizeof(float4) * (size_t)glo}, {e->pid[nxt] - glo, sizeof(int) * (size_t)glo}};
        BbxSeg shi[4] = {{e->cell_start[nxt] + g.c_own1 - g.plane, tb}, {e->pos[nxt] + (n_new - nl), sizeof(float4) * (size_t)nl}, {e->vel[nxt] + (n_new - nl), sizeof(float4) * (size_t)nl}, {e->pid[nxt] + (n_new - nl), sizeof(int) * (size_t)nl}};
        BbxSeg rhi[4] = {{e->gtab + g.plane + 1, tb}, {e->pos[nxt] + n_new, sizeof(float4) * (size_t)ghi}, {e->vel[nxt] + n_new, sizeof(float4) * (size_t)ghi}, {e->pid[nxt] + n_new, sizeof(int) * (size_t)ghi}};
        COMM(e->comm->exchange(e->stream, slo, rlo, 4, shi, rhi, 4));
        e->n = n_new; e->n_first = nf; e->n_last = nl; e->n_glo = glo; e->n_ghi = ghi;
        e->n_hint = n_new; e->n_launch = std::max(e->n_launch, bound_of(e, n_new));
        // a particle that left the slab's halo (dropped by the hash kernel) lets the collective sub-step finish first -- every
        // exchange the neighbours wait for still happens -- and is reported by the stepping call on its way out
        if(e->st_host->error == BBX_ERR_OUT_OF_DOMAIN) e->deferred_error = BBX_ERR_OUT_OF_DOMAIN;
        else if(e->st_host->error) return set_error(e->st_host->error, "device-side error %d (%s)", e->st_host->error, device_error_text(e->st_host->error));
    }
    LAUNCH(e, k_ghost_table, div_up(g.plane, 256), 256, g, e->st, e->has_lo, e->has_hi, e->gtab, e->gtab + g.plane + 1, e->cell_start[nxt], e->cell[nxt], e->count);
    CU(cudaGetLastError());
    return BBX_OK;
}

// UpdateGridDistributionGPU minus the bucket fill (sph_equations3.cpp:511-539)
#define BBX_SMALL_GRID (e->sm_count * 4)
#define BBX_CHECK_GRID (e->sm_count) // grid of the full-rebuild kernels when they are only a flag check (they stride over the data if it fires)
static int grid_update(bbx_engine *e){
    const bool slab = IS_SLAB(e);
    if(e->n == 0 && !slab) return BBX_OK;
    if(slab && !e->peers_ready){ int rc = setup_peers(e); if(rc) return rc; }
    DevGrid &g = e->grid;
    const int cur = e->cur, nxt = cur ^ 1;
    if(slab){ int rc = halo_settle_all(e); if(rc) return rc; } // (the neighbours' x, v of the last integration: hashed below)
    // single domain: the host knows the counts.  Slab engines: they live in DevState (n_all = -1 tells the kernels to read
    // them there) and the launches cover the capacity -- nothing below waits for the host when the halo push is on.
    const int n_all = slab ? -1 : e->n, n_lo = slab ? 0 : 0, n_own = slab ? 0 : e->n;
    int n_bound = e->n;
    if(slab){
        // the slots to hash are those of the LAST update (covered by the bound it was checked against); the kernels behind
        // the scan work on the new count, covered by the bound chosen now from the newest count the host has seen
        { int rc = poll_hint(e); if(rc) return rc; }
        const int prev = e->n_launch;
        e->n_launch = e->p2p ? bound_of(e, std::max(e->n_hint, 1)) : bound_of(e, e->n);
        n_bound = std::max(prev, e->n_launch) + 2 * e->gcap;
    }
    int force = (e->force_full || !e->have_chains) ? 1 : 0;
    int par = e->epoch & 1;
    const int own_cells = g.c_own1 - g.c_own0;
    if(!force) CU(cudaMemsetAsync(e->movemask, 0, sizeof(unsigned) * (size_t)g.total, e->stream));
    LAUNCH(e, k_hash_count, div_up(std::max(std::max(n_bound, e->scan_tiles), 1), 256), 256, n_all, n_lo, n_own, e->pos[cur], e->cell[cur], e->newcell, e->count, g, e->st,
           force ? 0 : 1, par, e->scan_status, e->scan_tiles, e->movemask);
    // the big-move rule and the jump detection are global decisions (the reference rebuilds ALL chains): the
    // flags are reduced over the ranks on a side stream while the scan and the (speculative) fill run
    if(slab){
        CU(cudaEventRecord(e->ev_fork, e->stream));
        CU(cudaStreamWaitEvent(e->side, e->ev_fork, 0));
        COMM(e->comm->allreduce_max_u32(e->side, (unsigned *)e->st, 4)); // rebuild_flag[2], jump_flag[2]
        CU(cudaEventRecord(e->ev_join, e->side));
    }
    LAUNCH(e, k_scan_cells, e->scan_tiles, 256, e->count + g.c_own0, own_cells, g.c_own0, e->scan_status, e->st, e->cell_start[nxt] + g.c_own0, e->occ_cells);
    unsigned counts_seq = 0;
    if(slab && e->p2p) counts_seq = post_counts(e, nxt);   // the new table is complete: the sizes can travel while the fill runs
    if(!force){
        // persistent grid: 8 lanes per occupied cell, grid-stride over the compact list of occupied cells
        int groups = std::max(1, std::min(n_bound, own_cells));
        int blocks = std::min(div_up((long long)groups * 8, 256), e->sm_count * 8);
        LAUNCH(e, k_fill_incremental, blocks, 256, g, e->st, par, slab ? 1 : 0, e->occ_cells, e->cell_start[cur], e->cell_start[nxt], e->newcell,
               e->pos[cur], e->vel[cur], e->pid[cur], e->pos[nxt], e->vel[nxt], e->pid[nxt], e->cell[nxt], e->rec, e->movemask);
    }
    if(slab) CU(cudaStreamWaitEvent(e->stream, e->ev_join, 0));
    // full path (forced, or selected on the device by the big-move / jump flags); small grids when it is
    // only a flag check
    int fb_n = force ? div_up(std::max(n_bound, 1), 256) : std::min(div_up(std::max(n_bound, 1), 256), BBX_CHECK_GRID);
    int fb_c = force ? div_up(own_cells, 256) : std::min(div_up(own_cells, 256), BBX_CHECK_GRID);
    LAUNCH(e, k_full_scatter, fb_n, 256, n_all, n_lo, g, e->st, par, force, e->newcell, e->cell_start[nxt], e->count, e->perm);
    LAUNCH(e, k_full_sort_cells, fb_c, 256, g, e->st, par, force, e->cell_start[nxt], e->pid[cur], e->perm, e->count, -1, 0);
    LAUNCH(e, k_full_gather, fb_n, 256, e->st, par, force, e->perm, e->newcell, e->pos[cur], e->vel[cur], e->pid[cur],
           e->pos[nxt], e->vel[nxt], e->pid[nxt], e->cell[nxt], e->rec);
    CU(cudaGetLastError());
    if(slab){ int rc = slab_refresh_ghosts(e, nxt, counts_seq); if(rc) return rc; }
    e->cur = nxt;
    e->have_chains = 1;
    e->last_force = force;
    e->epoch++;
    return BBX_OK;
}

#ifdef BBX_LISTS_V7
// persistent grid of the list build: one warp per occupied cell, grid-stride over the occupied-cell list
static int list_blocks(bbx_engine *e){
    long long cells = std::min<long long>(launch_n(e), e->grid.c_own1 - e->grid.c_own0);
    return (int)std::max<long long>(1, std::min<long long>((cells + BBX_LW - 1) / BBX_LW, (long long)e->sm_count * e->list_ctas_per_sm));
}
#else
// persistent grid of the list build: one warp per tile of 32 consecutive slots, grid-stride over the tiles
static int list_blocks(bbx_engine *e){
    long long tiles = ((long long)launch_n(e) + 31) / 32;
    return (int)std::max<long long>(1, std::min<long long>((tiles + BBX_TP_WARPS - 1) / BBX_TP_WARPS, (long long)e->sm_count * e->list_ctas_per_sm));
}
#endif
static int phase_density(bbx_engine *e, const StepParams &P, int sph){
    int cur = e->cur;
    if(launch_n(e) > 0){
#ifndef BBX_LISTS_V7
        if(sph) LAUNCH_S(e, k_lists_density_tp<1>, list_blocks(e), BBX_TP_WARPS * 32, BBX_TP_SMEM, P, e->grid, e->st, e->cell[cur], e->pos[cur], e->vel[cur], e->cell_start[cur], e->nbr, e->nbr_cnt, e->pressure, e->posq, e->rec, halo_none());
        else if(e->overlap){
            // boundary tiles (their rho / records go to the neighbours, then the density flag) beside all the others
            int rc = ov_begin(e, HALO_DENSITY); if(rc) return rc;
            const int bb_ = std::min(div_up(boundary_blocks(e, 32), BBX_TP_WARPS), e->sm_count * e->list_ctas_per_sm);
            LAUNCH_ON(e, e->bstream, k_lists_density_tp<0>, bb_, BBX_TP_WARPS * 32, BBX_TP_SMEM, with_part(P, 1), e->grid, e->st, e->cell[cur], e->pos[cur], e->vel[cur], e->cell_start[cur], e->nbr, e->nbr_cnt, e->pressure, e->posq, e->rec,
                     halo_dst(e, e->peer[0].rec, e->peer[1].rec));
            halo_signal(e, HALO_DENSITY, e->bstream);
            LAUNCH_S(e, k_lists_density_tp<0>, list_blocks(e), BBX_TP_WARPS * 32, BBX_TP_SMEM, with_part(P, 2), e->grid, e->st, e->cell[cur], e->pos[cur], e->vel[cur], e->cell_start[cur], e->nbr, e->nbr_cnt, e->pressure, e->posq, e->rec,
                     halo_none());
            CU(cudaGetLastError());
            return ov_end(e, HALO_DENSITY);
        }
        else LAUNCH_S(e, k_lists_density_tp<0>, list_blocks(e), BBX_TP_WARPS * 32, BBX_TP_SMEM, P, e->grid, e->st, e->cell[cur], e->pos[cur], e->vel[cur], e->cell_start[cur], e->nbr, e->nbr_cnt, e->pressure, e->posq, e->rec,
                    halo_dst(e, e->peer[0].rec, e->peer[1].rec));
#else
        if(sph) LAUNCH(e, k_cell_lists_density<1>, list_blocks(e), BBX_LT, P, e->grid, e->st, e->occ_cells, e->pos[cur], e->vel[cur], e->cell_start[cur], e->nbr, e->nbr_cnt, e->pressure, e->posq, e->rec, halo_none());
        else LAUNCH(e, k_cell_lists_density<0>, list_blocks(e), BBX_LT, P, e->grid, e->st, e->occ_cells, e->pos[cur], e->vel[cur], e->cell_start[cur], e->nbr, e->nbr_cnt, e->pressure, e->posq, e->rec,
                    halo_dst(e, e->peer[0].rec, e->peer[1].rec));
#endif
        CU(cudaGetLastError());
    }
    if(!sph && e->overlap){   // (a slab without particles still signals)
        int rc = ov_begin(e, HALO_DENSITY); if(rc) return rc;
        halo_signal(e, HALO_DENSITY, e->bstream);
        return ov_end(e, HALO_DENSITY);
    }
    if(!sph && e->p2p) return halo_sync(e, HALO_DENSITY);
    // ghost rho (and, for the SPH step, p / rho^2): both force sweeps read the 32-byte records (x, rho | v, p / rho^2)
    { void *arr[1] = {e->rec}; size_t z[1] = {2 * sizeof(float4)}; return exchange_planes(e, arr, z, 1); }
}
// grid of a list sweep: one thread per particle
static int sweep_grid(bbx_engine *e){ return div_up(launch_n(e), BBX_BS); }
// grid of a staged sweep: one CTA per tile of BBX_TS consecutive slots
static int tile_grid(bbx_engine *e){ return div_up(launch_n(e), BBX_TS); }
static int phase_force_np_predict(bbx_engine *e, const StepParams &P){
    int cur = e->cur;
    if(e->overlap){
        int rc = ov_begin(e, HALO_PREDICT); if(rc) return rc;              // (the boundary blocks read the neighbours' rho)
        for(int part = 1; part <= 2; part++){
            cudaStream_t strm = part == 1 ? e->bstream : e->stream;
            if(launch_n(e) > 0){
                const StepParams Q = with_part(P, part);
                int *q = part == 1 ? e->queue_b : e->queue;
                const HaloDst H = part == 1 ? halo_dst(e, e->peer[0].pred, e->peer[1].pred) : halo_none();
                LAUNCH_ON(e, strm, k_force_np_predict, part == 1 ? boundary_blocks(e, BBX_BS) : sweep_grid(e), BBX_BS, 0, Q, e->grid, e->st, e->cull, e->pos[cur], e->vel[cur], e->rec, e->cell[cur], e->cell_start[cur],
                       e->nbr, e->nbr_cnt, e->force, e->pred, q, H);
                LAUNCH_ON(e, strm, k_collide_predict, BBX_SMALL_GRID, 128, 0, Q, e->st, e->colliders, q, e->pos[cur], e->vel[cur], e->force, e->pred, H);
                CU(cudaGetLastError());
            }
            if(part == 1) halo_signal(e, HALO_PREDICT, e->bstream);
        }
        return ov_end(e, HALO_PREDICT);
    }
    if(launch_n(e) > 0){
        LAUNCH(e, k_force_np_predict, sweep_grid(e), BBX_BS, P, e->grid, e->st, e->cull, e->pos[cur], e->vel[cur], e->rec, e->cell[cur], e->cell_start[cur],
               e->nbr, e->nbr_cnt, e->force, e->pred, e->queue, halo_dst(e, e->peer[0].pred, e->peer[1].pred));
        LAUNCH(e, k_collide_predict, BBX_SMALL_GRID, 128, P, e->st, e->colliders, e->queue, e->pos[cur], e->vel[cur], e->force, e->pred,
               halo_dst(e, e->peer[0].pred, e->peer[1].pred));
        CU(cudaGetLastError());
    }
    if(e->p2p) return halo_sync(e, HALO_PREDICT);
    return exchange1(e, e->pred); // ghost x*
}
static int phase_pressure(bbx_engine *e, const StepParams &P, int first){
    int cur = e->cur;
    if(e->overlap){
        int rc = ov_begin(e, HALO_PRESSURE); if(rc) return rc;             // (the boundary blocks read the neighbours' x*)
        for(int part = 1; part <= 2; part++){
            if(launch_n(e) > 0){
                LAUNCH_ON(e, part == 1 ? e->bstream : e->stream, k_pressure, part == 1 ? boundary_blocks(e, BBX_TS) : tile_grid(e), BBX_TS, BBX_STAGE_BYTES(3), with_part(P, part), e->grid, e->st, first, e->pos[cur], e->pred, e->cell[cur], e->cell_start[cur],
                         e->nbr, e->nbr_cnt, e->pressure, e->rho_pred, e->rho_err, e->posq, part == 1 ? halo_dst(e, e->peer[0].posq, e->peer[1].posq) : halo_none());
                CU(cudaGetLastError());
            }
            if(part == 1) halo_signal(e, HALO_PRESSURE, e->bstream);
        }
        return ov_end(e, HALO_PRESSURE);
    }
    if(launch_n(e) > 0){
        LAUNCH_S(e, k_pressure, tile_grid(e), BBX_TS, BBX_STAGE_BYTES(3), P, e->grid, e->st, first, e->pos[cur], e->pred, e->cell[cur], e->cell_start[cur],
               e->nbr, e->nbr_cnt, e->pressure, e->rho_pred, e->rho_err, e->posq, halo_dst(e, e->peer[0].posq, e->peer[1].posq));
        CU(cudaGetLastError());
    }
    if(e->p2p) return halo_sync(e, HALO_PRESSURE);
    return exchange1(e, e->posq); // ghost (x, p / rho*^2)
}
static int phase_pressure_force(bbx_engine *e, const StepParams &P, int integrate){
    int cur = e->cur; int nb = sweep_grid(e);
    if(e->overlap && !integrate){   // ("correct" mode: one launch, after everything the pressure sweep produced, the neighbours' part included)
        CU(cudaStreamWaitEvent(e->stream, e->ev_b[HALO_PRESSURE], 0));
        int rc = halo_settle(e, HALO_PRESSURE); if(rc) return rc;
    }
    if(e->overlap && integrate){
        int rc = ov_begin(e, HALO_INTEGRATE); if(rc) return rc;            // (the boundary blocks read the neighbours' (x, p / rho*^2))
        for(int part = 1; part <= 2; part++){
            cudaStream_t strm = part == 1 ? e->bstream : e->stream;
            if(launch_n(e) > 0){
                const StepParams Q = with_part(P, part);
                int *q = part == 1 ? e->queue_b : e->queue;
                const HaloDst H = part == 1 ? halo_dst(e, e->peer[0].pos[cur], e->peer[1].pos[cur], e->peer[0].vel[cur], e->peer[1].vel[cur]) : halo_none();
                LAUNCH_ON(e, strm, k_pressure_force<1>, part == 1 ? boundary_blocks(e, BBX_BS) : nb, BBX_BS, 0, Q, e->grid, e->st, e->cull, e->pos[cur], e->vel[cur], e->posq, e->cell[cur], e->cell_start[cur], e->nbr, e->nbr_cnt, e->force, e->force_p, q, H);
                LAUNCH_ON(e, strm, k_collide_integrate, BBX_SMALL_GRID, 128, 0, Q, e->grid, e->st, e->colliders, q, e->pos[cur], e->vel[cur], e->force, H);
                CU(cudaGetLastError());
            }
            if(part == 1) halo_signal(e, HALO_INTEGRATE, e->bstream);   // (waited for by the next grid update)
        }
        return ov_end(e, HALO_INTEGRATE);
    }
    if(launch_n(e) > 0){
        if(integrate){
            LAUNCH(e, k_pressure_force<1>, nb, BBX_BS, P, e->grid, e->st, e->cull, e->pos[cur], e->vel[cur], e->posq, e->cell[cur], e->cell_start[cur], e->nbr, e->nbr_cnt, e->force, e->force_p, e->queue,
                   halo_dst(e, e->peer[0].pos[cur], e->peer[1].pos[cur], e->peer[0].vel[cur], e->peer[1].vel[cur]));
            LAUNCH(e, k_collide_integrate, BBX_SMALL_GRID, 128, P, e->grid, e->st, e->colliders, e->queue, e->pos[cur], e->vel[cur], e->force,
                   halo_dst(e, e->peer[0].pos[cur], e->peer[1].pos[cur], e->peer[0].vel[cur], e->peer[1].vel[cur]));
        }else{
            LAUNCH(e, k_pressure_force<0>, nb, BBX_BS, P, e->grid, e->st, e->cull, e->pos[cur], e->vel[cur], e->posq, e->cell[cur], e->cell_start[cur], e->nbr, e->nbr_cnt, e->force, e->force_p, e->queue, halo_none());
        }
        CU(cudaGetLastError());
    }
    if(integrate && e->p2p) return halo_sync(e, HALO_INTEGRATE);
    // after the integration the neighbours need the new x, v of my boundary planes (in the old order) to
    // run their own grid update: that is where migrating particles change owner
    if(integrate) return exchange2(e, e->pos[cur], e->vel[cur]);
    return BBX_OK;
}
static int phase_integrate(bbx_engine *e, const StepParams &P, int with_fp){
    int cur = e->cur;
    if(launch_n(e) > 0){
        LAUNCH(e, k_integrate, div_up(launch_n(e), 256), 256, P, e->grid, e->st, e->colliders, e->cull, e->pos[cur], e->vel[cur], e->force, with_fp ? e->force_p : (const float4 *)nullptr);
        CU(cudaGetLastError());
    }
    return exchange2(e, e->pos[cur], e->vel[cur]);
}
static int phase_pseudo_viscosity(bbx_engine *e, const StepParams &P, double dt){
    if(!(e->cfg.pseudo_viscosity * dt > 0.1)) return BBX_OK; // sph_equations3.cpp:655-657
    int cur = e->cur;
    // (the ghosts' rho for the aggregation weights rides in vel.w: it came with the post-integration halo)
    if(launch_n(e) > 0){
        LAUNCH(e, k_pseudo_aggregate, sweep_grid(e), BBX_BS, P, e->grid, e->pos[cur], e->vel[cur], e->cell[cur], e->cell_start[cur], e->nbr, e->nbr_cnt, e->smoothed);
        LAUNCH(e, k_pseudo_interpolate, div_up(launch_n(e), 256), 256, P, e->vel[cur], e->smoothed);
        CU(cudaGetLastError());
    }
    return exchange1(e, e->vel[cur]);
}
static int read_state(bbx_engine *e){
    CU(cudaMemcpyAsync(e->st_host, e->st, sizeof(DevState), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return BBX_OK;
}

static int step_pcisph(bbx_engine *e, double dt){
    if(e->n == 0 && !IS_SLAB(e)) return BBX_OK;
    if(!(dt > 0)) return set_error(BBX_ERR_INVALID, "dt must be positive");
    StepParams P;
    int rc;
    if((rc = entry_error(e))) return rc;
    tick(e, T_GRID);
    if((rc = grid_update(e))) return rc;
    make_params(e, dt, P);
    tick(e, T_DENSITY);
    if((rc = phase_density(e, P, 0))) return rc;
    tick(e, T_FORCE_NP);
    if((rc = phase_force_np_predict(e, P))) return rc;
    if(e->cfg.pcisph_reference_compat){
        // the reference's loop always leaves after one iteration (density error never stored)
        tick(e, T_PRESSURE);
        if((rc = phase_pressure(e, P, 1))) return rc;
        tick(e, T_PRESSURE_FORCE);
        if((rc = phase_pressure_force(e, P, 1))) return rc;
        e->st_host->iterations = 1;
    }else{
        int it = 0;
        for(int k = 0; k < e->cfg.pcisph_max_iterations; k++){
            if(k > 0){
                tick(e, T_PREDICT);
                if(launch_n(e) > 0) LAUNCH(e, k_predict_again, div_up(launch_n(e), 256), 256, P, e->colliders, e->cull, e->pos[e->cur], e->vel[e->cur], e->force, e->force_p, e->pred);
                if((rc = exchange1(e, e->pred))) return rc;
            }
            tick(e, T_PRESSURE);
            if((rc = phase_pressure(e, P, k == 0))) return rc;
            tick(e, T_PRESSURE_FORCE);
            if((rc = phase_pressure_force(e, P, 0))) return rc;
            it++;
            // the loop exit test needs max |rho* - rho0| on the host, as in the reference (pcisph_equations3.cpp:238-246)
            if(IS_SLAB(e)) COMM(e->comm->allreduce_max_u32(e->stream, &e->st->max_err_bits, 1));
            if((rc = read_state(e))) return rc;
            float maxerr; memcpy(&maxerr, &e->st_host->max_err_bits, 4);
            if(fabs((double)maxerr / e->cfg.target_density) < e->cfg.pcisph_max_density_error_ratio) break;
            if(k + 1 < e->cfg.pcisph_max_iterations){ CU(cudaMemsetAsync(&e->st->max_err_bits, 0, 4, e->stream)); }
            // (the queue of the predict pre-check is only used on the first iteration)
        }
        tick(e, T_INTEGRATE);
        if((rc = phase_integrate(e, P, 1))) return rc;
        e->st_host->iterations = it;
    }
    if((rc = phase_pseudo_viscosity(e, P, dt))) return rc;
    tick(e, T_COUNT);
    probe_error(e);
    e->substeps++;
    if(e->comm && e->comm->shares_device()) CU(cudaStreamSynchronize(e->stream));
    if(e->deferred_error) return set_error(e->deferred_error, "device-side error %d (%s)", e->deferred_error, device_error_text(e->deferred_error));
    return BBX_OK;
}

static int step_sph(bbx_engine *e, double dt){
    if(e->n == 0 && !IS_SLAB(e)) return BBX_OK;
    if(!(dt > 0)) return set_error(BBX_ERR_INVALID, "dt must be positive");
    StepParams P;
    int rc;
    if((rc = entry_error(e))) return rc;
    tick(e, T_GRID);
    if((rc = grid_update(e))) return rc;
    make_params(e, dt, P);
    tick(e, T_DENSITY);
    if((rc = phase_density(e, P, 1))) return rc;
    tick(e, T_FORCE_NP);
    if(launch_n(e) > 0){
        LAUNCH(e, k_sph_forces, sweep_grid(e), BBX_BS, P, e->grid, e->posq, e->vel[e->cur], e->rec, e->cell[e->cur], e->cell_start[e->cur], e->nbr, e->nbr_cnt, e->force);
        CU(cudaGetLastError());
    }
    tick(e, T_INTEGRATE);
    if((rc = phase_integrate(e, P, 0))) return rc;
    if((rc = phase_pseudo_viscosity(e, P, dt))) return rc;
    tick(e, T_COUNT);
    probe_error(e);
    e->substeps++;
    if(e->comm && e->comm->shares_device()) CU(cudaStreamSynchronize(e->stream));
    if(e->deferred_error) return set_error(e->deferred_error, "device-side error %d (%s)", e->deferred_error, device_error_text(e->deferred_error));
    return BBX_OK;
}

int bbx_step_pcisph(bbx_engine *e, double dt){ CHECK_ENGINE(e); int rc = step_pcisph(e, dt); if(rc) return rc; if(e->timing) harvest(e); return BBX_OK; }
int bbx_step_sph(bbx_engine *e, double dt){ CHECK_ENGINE(e); int rc = step_sph(e, dt); if(rc) return rc; if(e->timing) harvest(e); return BBX_OK; }

int bbx_step_many(bbx_engine *e, double dt, int solver, int n){
    CHECK_ENGINE(e);
    for(int s = 0; s < n; s++){
        int rc = solver == BBX_SOLVER_SPH ? step_sph(e, dt) : step_pcisph(e, dt);
        if(rc) return rc;
        if(e->timing && e->ev_used > 3000) harvest(e);
    }
    if(e->timing) harvest(e);
    return BBX_OK;
}
// n sub-steps bracketed by two events on the engine's stream: *ms = device time of the whole region (launch gaps
// included), the figure bench.py reports; returns after the last sub-step has finished
int bbx_step_many_timed(bbx_engine *e, double dt, int solver, int n, float *ms){
    CHECK_ENGINE(e);
    if(!ms) return set_error(BBX_ERR_INVALID, "null");
    cudaEvent_t e0, e1; CU(cudaEventCreate(&e0)); if(cudaEventCreate(&e1) != cudaSuccess){ cudaEventDestroy(e0); return set_error(BBX_ERR_CUDA, "cudaEventCreate failed"); }
    int rc = BBX_OK;
    if(cudaEventRecord(e0, e->stream) != cudaSuccess) rc = set_error(BBX_ERR_CUDA, "cudaEventRecord failed");
    for(int s = 0; s < n && rc == BBX_OK; s++) rc = solver == BBX_SOLVER_SPH ? step_sph(e, dt) : step_pcisph(e, dt);
    if(rc == BBX_OK && (cudaEventRecord(e1, e->stream) != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess)) rc = set_error(BBX_ERR_CUDA, "event synchronisation failed: %s", cudaGetErrorString(cudaGetLastError()));
    if(rc == BBX_OK){ *ms = 0.f; cudaEventElapsedTime(ms, e0, e1); rc = sticky_error(e); }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if(e->timing) harvest(e);
    return rc;
}

int bbx_run_phase(bbx_engine *e, int phase, double dt){
    CHECK_ENGINE(e);
    if(e->n == 0 && !IS_SLAB(e)) return BBX_OK;
    StepParams P; make_params(e, dt, P);
    switch(phase){
        case BBX_PHASE_GRID: return grid_update(e);
        case BBX_PHASE_DENSITY: { int rc = phase_density(e, P, 0); return rc ? rc : ov_join(e); }
        case BBX_PHASE_FORCE_NP: { int rc = phase_force_np_predict(e, P); return rc ? rc : ov_join(e); }
        case BBX_PHASE_PREDICT: return BBX_OK; // fused into FORCE_NP on the first iteration
        case BBX_PHASE_PRESSURE: { int rc = phase_pressure(e, P, 1); return rc ? rc : ov_join(e); }
        case BBX_PHASE_PRESSURE_FORCE: return phase_pressure_force(e, P, 0);
        case BBX_PHASE_INTEGRATE: { int rc = phase_integrate(e, P, 1); if(rc) return rc; e->substeps++; return phase_pseudo_viscosity(e, P, dt); }
    }
    return set_error(BBX_ERR_INVALID, "unknown phase %d", phase);
}

// SphParticleSet3::ComputeNumberOfTimeSteps (src/core/particle.h:584-608) with the device-reduced max |f|
static unsigned number_of_time_steps(bbx_engine *e, double remaining, double max_force, double scale){
    double limit = 0.40 * e->h / e->cfg.sound_speed;
    if(!(fabs(max_force) < 1e-8)){
        double by_force = 0.25 * sqrt(e->h * e->mass / max_force);
        limit = std::min(by_force, limit);
    }
    return (unsigned)ceil(remaining / (scale * limit));
}

int bbx_advance(bbx_engine *e, double seconds, int solver, int *substeps, float *ms){
    CHECK_ENGINE(e);
    cudaEvent_t e0, e1; CU(cudaEventCreate(&e0)); if(cudaEventCreate(&e1) != cudaSuccess){ cudaEventDestroy(e0); return set_error(BBX_ERR_CUDA, "cudaEventCreate failed"); }
    int rc = BBX_OK;
    if(cudaEventRecord(e0, e->stream) != cudaSuccess) rc = set_error(BBX_ERR_CUDA, "cudaEventRecord failed");
    double remaining = seconds; int count = 0;
    double scale = solver == BBX_SOLVER_SPH ? 1.0 : e->cfg.time_step_limit_scale;
    while(rc == BBX_OK && remaining > (double)0.0001f){ // `while(remainingTime > Epsilon)`, pcisph_solver3.cpp:76
        if(IS_SLAB(e) && e->comm->allreduce_max_u32(e->stream, &e->st->max_force_bits, 1)){ rc = set_error(BBX_ERR_COMM, "%s", e->comm->err.c_str()); break; } // same dt on every slab
        if((rc = read_state(e))) break;
        // the CFL scan synchronises anyway: a device-side failure of the previous sub-step ends the frame here
        if(e->st_host->error){ rc = set_error(e->st_host->error, "device-side error %d (%s)", e->st_host->error, device_error_text(e->st_host->error)); break; }
        float mf; memcpy(&mf, &e->st_host->max_force_bits, 4);
        unsigned nsteps = number_of_time_steps(e, remaining, (double)mf, scale);
        double dt = remaining / (double)nsteps;
        rc = solver == BBX_SOLVER_SPH ? step_sph(e, dt) : step_pcisph(e, dt);
        if(rc) break;
        remaining -= dt; count++;
    }
    if(rc == BBX_OK && (cudaEventRecord(e1, e->stream) != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess)) rc = set_error(BBX_ERR_CUDA, "event synchronisation failed");
    float t = 0.f; if(rc == BBX_OK) cudaEventElapsedTime(&t, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if(rc == BBX_OK) rc = sticky_error(e);
    if(rc) return rc;
    if(e->timing) harvest(e);
    if(substeps) *substeps = count;
    if(ms) *ms = t;
    return BBX_OK;
}

int bbx_synchronize(bbx_engine *e){ CHECK_ENGINE(e); CU(cudaStreamSynchronize(e->stream)); return sticky_error(e); }
int bbx_set_timing(bbx_engine *e, int enabled){ CHECK_ENGINE(e); if(e->timing) harvest(e); e->timing = enabled ? 1 : 0; e->ev_used = 0; return BBX_OK; }

int bbx_stats(bbx_engine *e, bbx_step_stats *out){
    CHECK_ENGINE(e);
    if(!out) return set_error(BBX_ERR_INVALID, "null");
    int it = e->st_host->iterations;
    int rc = read_state(e); if(rc) return rc;
    e->st_host->iterations = it;
    if(e->counts_stale){ e->counts_stale = 1; rc = sync_counts(e); if(rc) return rc; e->st_host->iterations = it; }
    const DevState &s = *e->st_host;
    memset(out, 0, sizeof(*out));
    out->particles = e->n; out->ghosts = e->n_glo + e->n_ghi; out->substeps = e->substeps; out->pcisph_iterations = it;
    const int done = (e->epoch + 1) & 1, next = e->epoch & 1; // flag slots of the last / the next grid update
    out->full_rebuild = (e->epoch > 0 && (e->last_force | s.rebuild_flag[done] | s.jump_flag[done])) ? 1 : 0;
    out->rebuild_flag = s.rebuild_flag[next]; out->neighbor_overflow = s.overflow;
    out->lost_particles = s.lost[done]; out->clamped = s.clamped; out->nan_count = s.nan_count;
    memcpy(&out->max_force, &s.max_force_bits, 4); memcpy(&out->max_density_error, &s.max_err_bits, 4);
    out->ms_grid = e->last_ms_grid; out->ms_step = e->last_ms_step;
    out->exact_passes = s.exact_passes; out->max_candidates = s.max_candidates; out->occupied_cells = s.n_occ; out->unstaged_tiles = s.unstaged_tiles;
    if(s.error) return set_error(s.error, "device-side error %d (%s)", s.error, device_error_text(s.error));
    return BBX_OK;
}

// ---------------------------------------------------------------------------------------- results
static int download(bbx_engine *e, int field, void *dst, int dtype, int compact){
    if(!dst) return set_error(BBX_ERR_INVALID, "null destination");
    { int rc_ = sync_counts(e); if(rc_) return rc_; }
    if(e->n == 0) return BBX_OK;
    int cur = e->cur; int n = e->n;
    const float4 *s4 = nullptr; const float *s1 = nullptr; const int *si = nullptr; int comps = 3;
    switch(field){
        case BBX_POSITION: s4 = e->pos[cur]; break;
        case BBX_VELOCITY: s4 = e->vel[cur]; break;
        case BBX_FORCE: case BBX_FORCE_NP: s4 = e->force; break;
        case BBX_PRED_POSITION: s4 = e->pred; break;
        case BBX_PRESSURE_FORCE: s4 = e->force_p; break;
        case BBX_DENSITY: s4 = e->vel[cur]; comps = 1; break;
        case BBX_PRESSURE: s1 = e->pressure; comps = 1; break;
        case BBX_PRED_DENSITY: s1 = e->rho_pred; comps = 1; break;
        case BBX_DENSITY_ERROR: s1 = e->rho_err; comps = 1; break;
        case BBX_NEIGHBOR_COUNT: si = e->nbr_cnt; comps = 1; break;
        default: return set_error(BBX_ERR_INVALID, "unknown field %d", field);
    }
    if(si){ if(dtype != BBX_I32) return set_error(BBX_ERR_INVALID, "field %d is BBX_I32", field); }
    else if(dtype != BBX_F32 && dtype != BBX_F64) return set_error(BBX_ERR_INVALID, "field %d needs BBX_F32 or BBX_F64", field);
    size_t esz = (dtype == BBX_F64) ? 8 : 4; size_t bytes = esz * comps * (size_t)n;
    int rc = ensure_stage(e, bytes); if(rc) return rc;
    LAUNCH(e, k_download, div_up(n, 256), 256, n, compact ? (const int *)nullptr : e->pid[cur], s4, s1, si, comps, dtype == BBX_F64, e->stage);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(dst, e->stage, bytes, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return BBX_OK;
}
int bbx_download(bbx_engine *e, int field, void *dst, int dtype){
    CHECK_ENGINE(e);
    if(IS_SLAB(e)) return set_error(BBX_ERR_INVALID, "bbx_download scatters by particle id over the whole set: use bbx_download_owned on a slab engine");
    return download(e, field, dst, dtype, 0);
}
int bbx_download_owned(bbx_engine *e, int field, void *dst, int dtype, int *ids, int *count){
    CHECK_ENGINE(e);
    { int rc_ = sync_counts(e); if(rc_) return rc_; }
    if(count) *count = e->n;
    if(e->n == 0) return BBX_OK;
    if(ids){
        CU(cudaMemcpyAsync(ids, e->pid[e->cur], sizeof(int) * (size_t)e->n, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
    }
    if(!dst) return BBX_OK;
    return download(e, field, dst, dtype, 1);
}

// positions + velocities (+ ids) of one frame with ONE kernel, back-to-back copies and ONE synchronisation: what a host
// run loop reads after Advance (UtilRunSimulation3, src/core/util.h:583-590)
int bbx_download_state(bbx_engine *e, void *pos, void *vel, int *ids, int dtype, int owned_order, int *count){
    CHECK_ENGINE(e);
    if(dtype != BBX_F32 && dtype != BBX_F64) return set_error(BBX_ERR_INVALID, "dtype must be BBX_F32 or BBX_F64");
    if(IS_SLAB(e) && !owned_order) return set_error(BBX_ERR_INVALID, "a slab engine returns its owned particles in cell order: pass owned_order = 1");
    { int rc_ = sync_counts(e); if(rc_) return rc_; }
    if(count) *count = e->n;
    if(e->n == 0) return BBX_OK;
    if(!pos || !vel) return set_error(BBX_ERR_INVALID, "null destination");
    const int n = e->n, cur = e->cur;
    const size_t bytes = (dtype == BBX_F64 ? 8 : 4) * 3 * (size_t)n;
    int rc = ensure_stage(e, 2 * bytes); if(rc) return rc;
    char *sp = (char *)e->stage, *sv = sp + bytes;
    LAUNCH(e, k_download_state, div_up(n, 256), 256, n, owned_order ? (const int *)nullptr : e->pid[cur], e->pos[cur], e->vel[cur], dtype == BBX_F64, sp, sv);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(pos, sp, bytes, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(vel, sv, bytes, cudaMemcpyDeviceToHost, e->stream));
    if(ids) CU(cudaMemcpyAsync(ids, e->pid[cur], sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return sticky_error(e);
}

int bbx_query_cells(bbx_engine *e, int n, const int *cells, const double *points, double d, int *cell_size, int *blocked){
    CHECK_ENGINE(e);
    if(n <= 0) return BBX_OK;
    if(!cells || !points || !cell_size || !blocked) return set_error(BBX_ERR_INVALID, "null");
    if(!e->have_chains){ for(int k = 0; k < n; k++){ cell_size[k] = 0; blocked[k] = 0; } return BBX_OK; }
    const size_t bi = sizeof(int) * (size_t)n, bp = sizeof(double) * 3 * (size_t)n;
    int rc = ensure_stage(e, bp + 3 * bi); if(rc) return rc;
    double *dp = (double *)e->stage; int *dc = (int *)((char *)e->stage + bp), *ds = dc + n, *db = ds + n;
    CU(cudaMemcpyAsync(dp, points, bp, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(dc, cells, bi, cudaMemcpyHostToDevice, e->stream));
    LAUNCH(e, k_query_cells, div_up(n, 256), 256, n, e->grid, dc, dp, d, e->cell_start[e->cur], e->pos[e->cur], ds, db);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(cell_size, ds, bi, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(blocked, db, bi, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return BBX_OK;
}

int bbx_export_cells(bbx_engine *e, int *cell_count, int *cell_order){
    CHECK_ENGINE(e);
    { int rc_ = sync_counts(e); if(rc_) return rc_; }
    if(!cell_count || !cell_order) return set_error(BBX_ERR_INVALID, "null");
    // cell_count covers the GLOBAL grid; a slab engine fills the cells it owns (others 0) and cell_order holds
    // its owned particles: the single-domain order is the concatenation of the slabs' orders
    const DevGrid &g = e->grid;
    size_t gtotal = (size_t)g.plane * g.gnz;
    memset(cell_count, 0, sizeof(int) * gtotal);
    if(!e->have_chains || e->n == 0) return BBX_OK;
    int own_cells = g.c_own1 - g.c_own0;
    int rc = ensure_stage(e, sizeof(int) * (size_t)own_cells); if(rc) return rc;
    LAUNCH(e, k_export_cells, div_up(own_cells, 256), 256, own_cells, e->cell_start[e->cur] + g.c_own0, (int *)e->stage);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(cell_count + (size_t)(g.zoff + g.own_z0) * g.plane, e->stage, sizeof(int) * (size_t)own_cells, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(cell_order, e->pid[e->cur], sizeof(int) * (size_t)e->n, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return BBX_OK;
}

static int export_neighbors(bbx_engine *e, int *counts, int *ids, int compact){
    if(!counts || !ids) return set_error(BBX_ERR_INVALID, "null");
    { int rc_ = sync_counts(e); if(rc_) return rc_; }
    if(e->n == 0) return BBX_OK;
    int n = e->n; size_t bytes = sizeof(int) * (size_t)n * (BBX_MAX_NEIGHBORS + 1);
    int rc = ensure_stage(e, bytes); if(rc) return rc;
    int *dc = (int *)e->stage; int *di = dc + n;
    LAUNCH(e, k_export_neighbors, div_up(n, BBX_BS), BBX_BS, n, e->grid, e->pid[e->cur], e->cell[e->cur], e->cell_start[e->cur], e->nbr, e->nbr_cnt, dc, di, compact);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(counts, dc, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(ids, di, sizeof(int) * (size_t)n * BBX_MAX_NEIGHBORS, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return BBX_OK;
}
int bbx_export_neighbors(bbx_engine *e, int *counts, int *ids){
    CHECK_ENGINE(e);
    if(IS_SLAB(e)) return set_error(BBX_ERR_INVALID, "use bbx_export_neighbors_owned on a slab engine");
    return export_neighbors(e, counts, ids, 0);
}
int bbx_export_neighbors_owned(bbx_engine *e, int *counts, int *ids){ CHECK_ENGINE(e); return export_neighbors(e, counts, ids, 1); }

int bbx_inject_chains(bbx_engine *e, const int *cell_count, const int *cell_order){
    CHECK_ENGINE(e);
    if(!cell_count || !cell_order) return set_error(BBX_ERR_INVALID, "null");
    if(IS_SLAB(e)) return set_error(BBX_ERR_INVALID, "bbx_inject_chains is not available on slab engines");
    if(e->n == 0) return BBX_OK;
    int total = e->grid.total, n = e->n, cur = e->cur, nxt = cur ^ 1;
    std::vector<int> start((size_t)total + 1), cell_of((size_t)n);
    long long acc = 0;
    for(int c = 0; c < total; c++){
        start[c] = (int)acc;
        if(cell_count[c] < 0) return set_error(BBX_ERR_INVALID, "negative chain length");
        for(int k = 0; k < cell_count[c]; k++){ if(acc + k >= n) return set_error(BBX_ERR_INVALID, "chains hold more than n particles"); cell_of[acc + k] = c; }
        acc += cell_count[c];
    }
    start[total] = (int)acc;
    if(acc != n) return set_error(BBX_ERR_INVALID, "chains hold %lld particles, engine has %d", acc, n);
    std::vector<char> seen((size_t)n, 0);
    for(int k = 0; k < n; k++){ int id = cell_order[k]; if(id < 0 || id >= n || seen[id]) return set_error(BBX_ERR_INVALID, "cell_order is not a permutation"); seen[id] = 1; }
    int rc = ensure_stage(e, sizeof(int) * 3 * (size_t)n); if(rc) return rc;
    int *d_order = (int *)e->stage, *d_slot = d_order + n, *d_cell = d_slot + n;
    CU(cudaMemcpyAsync(d_order, cell_order, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(d_cell, cell_of.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->cell_start[nxt], start.data(), sizeof(int) * ((size_t)total + 1), cudaMemcpyHostToDevice, e->stream));
    LAUNCH(e, k_slot_of_id, div_up(n, 256), 256, n, e->pid[cur], d_slot);
    LAUNCH(e, k_inject_gather, div_up(n, 256), 256, n, d_order, d_slot, d_cell, e->pos[cur], e->vel[cur], e->pos[nxt], e->vel[nxt], e->pid[nxt], e->cell[nxt], e->rec);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(e->stream)); // host vectors go out of scope
    e->cur = nxt; e->have_chains = 1; e->force_full = 0;
    return BBX_OK;
}

int bbx_set_rebuild_flag(bbx_engine *e, int flag){
    CHECK_ENGINE(e);
    int v = flag ? 1 : 0;
    CU(cudaMemcpyAsync(&e->st->rebuild_flag[e->epoch & 1], &v, sizeof(int), cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return BBX_OK;
}

int bbx_launch_count(bbx_engine *e, long long *count){ if(!e || !count) return set_error(BBX_ERR_INVALID, "null"); *count = e->launches; return BBX_OK; }
int bbx_kernel_time(bbx_engine *e, int phase, float *ms, int *launches){
    if(!e || phase < 0 || phase > T_COUNT) return set_error(BBX_ERR_INVALID, "bad phase");
    if(ms) *ms = e->phase_ms[phase];
    if(launches) *launches = e->phase_launches[phase];
    return BBX_OK;
}
int bbx_reset_kernel_time(bbx_engine *e){
    if(!e) return set_error(BBX_ERR_INVALID, "null");
    memset(e->phase_ms, 0, sizeof(e->phase_ms)); memset(e->phase_launches, 0, sizeof(e->phase_launches));
    return BBX_OK;
}

// ------------------------------------------------------------------------------------ slab groups
int bbx_comm_unique_id(unsigned char id[BBX_NCCL_ID_BYTES]){
    if(!id) return set_error(BBX_ERR_INVALID, "null");
    if(!g_nccl.load()) return set_error(BBX_ERR_COMM, "%s", g_nccl.load_error.c_str());
    ncclUniqueId uid;
    ncclResult_t r = g_nccl.GetUniqueId(&uid);
    if(r != ncclSuccess) return set_error(BBX_ERR_COMM, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
    static_assert(sizeof(ncclUniqueId) == BBX_NCCL_ID_BYTES, "ncclUniqueId size");
    memcpy(id, &uid, BBX_NCCL_ID_BYTES);
    return BBX_OK;
}
int bbx_comm_init(bbx_engine *e, int rank, int nranks, const unsigned char id[BBX_NCCL_ID_BYTES]){
    CHECK_ENGINE(e);
    if(!id || nranks < 1 || rank < 0 || rank >= nranks) return set_error(BBX_ERR_INVALID, "bad rank / nranks");
    if(e->comm) return set_error(BBX_ERR_INVALID, "engine already has a communicator");
    if((rank > 0) != (e->has_lo != 0) || (rank + 1 < nranks) != (e->has_hi != 0))
        return set_error(BBX_ERR_INVALID, "rank %d of %d does not match the slab [%d, %d) of %d planes (slabs must be ordered by z)", rank, nranks,
                         e->cfg.slab_z_begin, e->cfg.slab_z_end, e->grid.gnz);
    NcclComm *c = new NcclComm();
    if(c->init(rank, nranks, id)){ std::string m = c->err; delete c; return set_error(BBX_ERR_COMM, "%s", m.c_str()); }
    e->comm = c; e->peers_ready = 0; // the neighbours' arrays are mapped at the first collective call (grid update)
    return BBX_OK;
}
int bbx_halo_mode(bbx_engine *e, int *p2p){ if(!e || !p2p) return set_error(BBX_ERR_INVALID, "null"); *p2p = e->p2p; return BBX_OK; }
int bbx_comm_init_local(bbx_engine *e, int rank, int nranks, const char *group){
    CHECK_ENGINE(e);
    if(e->comm) return set_error(BBX_ERR_INVALID, "engine already has a communicator");
    if((rank > 0) != (e->has_lo != 0) || (rank + 1 < nranks) != (e->has_hi != 0))
        return set_error(BBX_ERR_INVALID, "rank %d of %d does not match the slab [%d, %d) of %d planes (slabs must be ordered by z)", rank, nranks,
                         e->cfg.slab_z_begin, e->cfg.slab_z_end, e->grid.gnz);
    LocalComm *c = new LocalComm();
    if(c->init(group, rank, nranks)){ std::string m = c->err; delete c; return set_error(BBX_ERR_COMM, "%s", m.c_str()); }
    e->comm = c; e->peers_ready = 0; // the neighbours' arrays are mapped at the first collective call (grid update)
    return BBX_OK;
}
// ParticleSet3 has no notion of slabs: this is the planner a multi-GPU host uses.  plane_counts[z] =
// particles per global cell plane; z_bounds[r] .. z_bounds[r+1] = planes of rank r, balanced by count with
// at least one plane per rank.
int bbx_slab_plan(int nplanes, const long long *plane_counts, int nranks, int *z_bounds){
    if(!plane_counts || !z_bounds || nranks < 1 || nplanes < nranks) return set_error(BBX_ERR_INVALID, "need at least one plane per rank");
    long long total = 0;
    for(int z = 0; z < nplanes; z++){ if(plane_counts[z] < 0) return set_error(BBX_ERR_INVALID, "negative plane count"); total += plane_counts[z]; }
    z_bounds[0] = 0;
    long long acc = 0; int z = 0;
    for(int r = 1; r < nranks; r++){
        // smallest z with prefix(z) >= r/nranks of the total, leaving a plane for every rank on both sides
        long long want = (total * r + nranks / 2) / nranks;
        int lo = z_bounds[r - 1] + 1, hi = nplanes - (nranks - r);
        while(z < lo){ acc += plane_counts[z]; z++; }
        while(z < hi && acc < want){ acc += plane_counts[z]; z++; }
        // step back one plane if that lands closer to the target
        if(z > lo && acc - want > want - (acc - plane_counts[z - 1])){ z--; acc -= plane_counts[z]; }
        z_bounds[r] = z;
    }
    z_bounds[nranks] = nplanes;
    return BBX_OK;
}
// one step from the cuts `current` towards `target` that bbx_rebalance can follow: every cut stays strictly inside the two
// slabs it separates (planes change hands between neighbours only, every rank keeps a plane); returns 1 in *done when the
// step reaches the target
int bbx_slab_plan_step(int nranks, const int *current, const int *target, int *step, int *done){
    if(!current || !target || !step || nranks < 1) return set_error(BBX_ERR_INVALID, "bad arguments");
    step[0] = current[0]; step[nranks] = current[nranks];
    int reached = 1;
    for(int r = 1; r < nranks; r++){
        int lo = current[r - 1] + 1, hi = current[r + 1] - 1;
        lo = std::max(lo, step[r - 1] + 1);                       // keep the step itself strictly increasing
        int c = std::min(std::max(target[r], lo), hi);
        step[r] = c;
        if(c != target[r]) reached = 0;
    }
    if(done) *done = reached;
    return BBX_OK;
}
// ---- re-balancing of the z cuts during a run (collective over the slab group, between two sub-steps)
// owned particles per GLOBAL cell plane, by the cells recorded at the last grid update (zero outside my planes)
int bbx_plane_counts(bbx_engine *e, long long *plane_counts){
    CHECK_ENGINE(e);
    if(!plane_counts) return set_error(BBX_ERR_INVALID, "null");
    const DevGrid &g = e->grid;
    for(int z = 0; z < g.gnz; z++) plane_counts[z] = 0;
    { int rc = sync_counts(e); if(rc) return rc; }
    if(e->n == 0 || !e->have_chains) return BBX_OK;
    const int planes = g.own_z1 - g.own_z0;
    std::vector<int> cs((size_t)planes + 1);
    CU(cudaMemcpy2DAsync(cs.data(), sizeof(int), e->cell_start[e->cur] + g.c_own0, sizeof(int) * (size_t)g.plane, sizeof(int), (size_t)planes + 1, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    for(int k = 0; k < planes; k++) plane_counts[g.zoff + g.own_z0 + k] = (long long)cs[k + 1] - cs[k];
    return BBX_OK;
}
// The slab group moves to the cuts z_bounds[0 .. nranks] (z_bounds[r] .. z_bounds[r + 1] = the planes of rank r; every rank
// passes the same array).  Whole planes change hands between NEIGHBOURS only -- a cut may move at most to the far end of
// the slab next to it, and every rank must keep at least one of its planes; call again to move further.  The particles of
// a plane travel as contiguous slot ranges (x, v, id, recorded cell) in their chain order, the receiver files them under
// the recorded cells with the re-sort of an append (old slots = chain order) and the boundary planes are exchanged again:
// cell orders, neighbour lists and the trajectory stay bit-identical to the single-domain engine.
int bbx_rebalance(bbx_engine *e, const int *z_bounds){
    CHECK_ENGINE(e);
    if(!IS_SLAB(e)) return BBX_OK;
    if(!z_bounds) return set_error(BBX_ERR_INVALID, "null");
    if(!e->comm) return set_error(BBX_ERR_INVALID, "slab engine without a communicator");
    if(!e->have_chains) return set_error(BBX_ERR_INVALID, "bbx_rebalance needs a particle set (bbx_set_particles_ids) first");
    int rc = halo_settle_all(e); if(rc) return rc;
    rc = sync_counts(e); if(rc) return rc;
    DevGrid &g = e->grid;
    const int rank = e->comm->rank, nranks = e->comm->nranks;
    const int z0 = g.zoff + g.own_z0, z1 = g.zoff + g.own_z1, nz0 = z_bounds[rank], nz1 = z_bounds[rank + 1];
    if(z_bounds[0] != 0 || z_bounds[nranks] != g.gnz) return set_error(BBX_ERR_INVALID, "z_bounds must run from 0 to the grid's %d planes", g.gnz);
    for(int r = 0; r < nranks; r++) if(z_bounds[r + 1] <= z_bounds[r]) return set_error(BBX_ERR_INVALID, "every rank needs at least one plane");
    const int k0 = std::max(z0, nz0), k1 = std::min(z1, nz1);      // the planes I keep
    // a plan this rank cannot follow must stop EVERY rank before the first exchange: agree on it (global max of a flag)
    const int cur = e->cur, nxt = cur ^ 1, n = e->n;
    {
        unsigned bad = (k1 <= k0) ? 1u : 0u;
        CU(cudaMemcpyAsync(e->perm, &bad, sizeof(unsigned), cudaMemcpyHostToDevice, e->stream));
        COMM(e->comm->allreduce_max_u32(e->stream, (unsigned *)e->perm, 1));
        unsigned any = 0;
        CU(cudaMemcpyAsync(&any, e->perm, sizeof(unsigned), cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        if(bad) return set_error(BBX_ERR_INVALID, "rank %d would keep none of its planes [%d, %d) in [%d, %d): move the cuts in smaller steps (bbx_slab_plan_step)", rank, z0, z1, nz0, nz1);
        if(any) return set_error(BBX_ERR_INVALID, "another rank of the slab group cannot follow this plan (a cut moved past a neighbouring slab)");
    }
    // slot boundaries of the planes that leave (cell table of the current buffer, owned part)
    int slot_k0 = 0, slot_k1 = n;
    if(k0 > z0) CU(cudaMemcpyAsync(&slot_k0, e->cell_start[cur] + g.c_own0 + (size_t)(k0 - z0) * g.plane, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    if(k1 < z1) CU(cudaMemcpyAsync(&slot_k1, e->cell_start[cur] + g.c_own0 + (size_t)(k1 - z0) * g.plane, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    const int s_lo = slot_k0, s_hi = n - slot_k1, n_keep = slot_k1 - slot_k0;
    const int to_lo[BBX_NCOUNTS] = {s_lo, 0}, to_hi[BBX_NCOUNTS] = {s_hi, 0};
    int from_lo[BBX_NCOUNTS] = {0, 0}, from_hi[BBX_NCOUNTS] = {0, 0};
    COMM(e->comm->neighbor_counts(e->stream, to_lo, to_hi, from_lo, from_hi));
    const int r_lo = e->has_lo ? from_lo[0] : 0, r_hi = e->has_hi ? from_hi[0] : 0;
    if((nz0 >= z0 && r_lo > 0) || (nz1 <= z1 && r_hi > 0)) return set_error(BBX_ERR_INVALID, "the ranks of the slab group disagree about z_bounds");
    const long long n_new = (long long)n_keep + r_lo + r_hi;
    {   // a rank that cannot hold its new share must stop EVERY rank before the exchange (agreed like the plan check above:
        // a lone early return would leave the neighbours waiting in the exchange)
        unsigned full = n_new > e->cap ? 1u : 0u, any = 0;
        CU(cudaMemcpyAsync(e->perm, &full, sizeof(unsigned), cudaMemcpyHostToDevice, e->stream));
        COMM(e->comm->allreduce_max_u32(e->stream, (unsigned *)e->perm, 1));
        CU(cudaMemcpyAsync(&any, e->perm, sizeof(unsigned), cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        if(full) return set_error(BBX_ERR_CAPACITY, "after re-balancing rank %d would own %lld particles, max_particles is %d", rank, n_new, e->cap);
        if(any) return set_error(BBX_ERR_CAPACITY, "another rank of the slab group cannot hold its share of this plan (its max_particles is too small)");
    }
    // the new grid of this rank (one ghost plane per neighbour, as at creation)
    const int zoff_new = nz0 - (e->has_lo ? 1 : 0);
    const long long plane = g.plane;
    // the other buffer is laid out [planes from below | kept | planes from above]: plane order = cell order, chains intact.
    // Kept particles: recorded cells re-based to the new local plane numbering
    if(n_keep > 0){
        CU(cudaMemcpyAsync(e->pos[nxt] + r_lo, e->pos[cur] + slot_k0, sizeof(float4) * (size_t)n_keep, cudaMemcpyDeviceToDevice, e->stream));
        CU(cudaMemcpyAsync(e->vel[nxt] + r_lo, e->vel[cur] + slot_k0, sizeof(float4) * (size_t)n_keep, cudaMemcpyDeviceToDevice, e->stream));
        CU(cudaMemcpyAsync(e->pid[nxt] + r_lo, e->pid[cur] + slot_k0, sizeof(int) * (size_t)n_keep, cudaMemcpyDeviceToDevice, e->stream));
        LAUNCH(e, k_cells_shift, div_up(n_keep, 256), 256, n_keep, e->cell[cur] + slot_k0, e->cell[nxt] + r_lo, (int)((long long)(g.zoff - zoff_new) * plane));
    }
    // the planes that leave travel with GLOBAL cell ids (staged in perm: s_lo + s_hi <= n slots)
    if(s_lo > 0) LAUNCH(e, k_cells_shift, div_up(s_lo, 256), 256, s_lo, e->cell[cur], e->perm, (int)((long long)g.zoff * plane));
    if(s_hi > 0) LAUNCH(e, k_cells_shift, div_up(s_hi, 256), 256, s_hi, e->cell[cur] + slot_k1, e->perm + s_lo, (int)((long long)g.zoff * plane));
    CU(cudaGetLastError());
    const size_t at_hi = (size_t)r_lo + (size_t)n_keep;
    {
        const size_t f4 = sizeof(float4), i4 = sizeof(int);
        BbxSeg slo[4] = {{e->pos[cur], f4 * (size_t)s_lo}, {e->vel[cur], f4 * (size_t)s_lo}, {e->pid[cur], i4 * (size_t)s_lo}, {e->perm, i4 * (size_t)s_lo}};
        BbxSeg shi[4] = {{e->pos[cur] + slot_k1, f4 * (size_t)s_hi}, {e->vel[cur] + slot_k1, f4 * (size_t)s_hi}, {e->pid[cur] + slot_k1, i4 * (size_t)s_hi}, {e->perm + s_lo, i4 * (size_t)s_hi}};
        BbxSeg rlo[4] = {{e->pos[nxt], f4 * (size_t)r_lo}, {e->vel[nxt], f4 * (size_t)r_lo}, {e->pid[nxt], i4 * (size_t)r_lo}, {e->cell[nxt], i4 * (size_t)r_lo}};
        BbxSeg rhi[4] = {{e->pos[nxt] + at_hi, f4 * (size_t)r_hi}, {e->vel[nxt] + at_hi, f4 * (size_t)r_hi}, {e->pid[nxt] + at_hi, i4 * (size_t)r_hi}, {e->cell[nxt] + at_hi, i4 * (size_t)r_hi}};
        COMM(e->comm->exchange(e->stream, slo, rlo, 4, shi, rhi, 4));
    }
    if(r_lo > 0) LAUNCH(e, k_cells_shift, div_up(r_lo, 256), 256, r_lo, e->cell[nxt], e->cell[nxt], (int)(-(long long)zoff_new * plane));
    if(r_hi > 0) LAUNCH(e, k_cells_shift, div_up(r_hi, 256), 256, r_hi, e->cell[nxt] + at_hi, e->cell[nxt] + at_hi, (int)(-(long long)zoff_new * plane));
    CU(cudaGetLastError());
    // switch to the new slab
    g.zoff = zoff_new;
    g.n[2] = (nz1 - nz0) + (e->has_lo ? 1 : 0) + (e->has_hi ? 1 : 0);
    g.own_z0 = e->has_lo ? 1 : 0; g.own_z1 = g.own_z0 + (nz1 - nz0);
    g.c_own0 = g.own_z0 * g.plane; g.c_own1 = g.own_z1 * g.plane;
    g.total = g.plane * g.n[2];
    e->cfg.slab_z_begin = nz0; e->cfg.slab_z_end = nz1;
    e->scan_tiles = div_up(g.c_own1 - g.c_own0, SCAN_TILE);
    e->n = (int)n_new;
    const int n_dev = (int)n_new;
    CU(cudaMemcpyAsync(&e->st->n_own, &n_dev, sizeof(int), cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream)); // (n_dev is a stack variable)
    e->n_hint = e->n; e->hint_count = 0; e->n_launch = bound_of(e, e->n);
    // the slots are in cell order: cell table + gather records from the recorded cells, then the boundary planes again
    LAUNCH(e, k_table_from_sorted, div_up(std::max(e->n, 1), 256), 256, e->n, g, e->cell[nxt], e->cell_start[nxt], e->pos[nxt], e->vel[nxt], e->rec);
    CU(cudaGetLastError());
    rc = slab_refresh_ghosts(e, nxt); if(rc) return rc;
    e->cur = nxt; e->have_chains = 1;
    return BBX_OK;
}

// global cell plane of each particle (the hash of Grid::GetHashedPosition, z component), host arithmetic
int bbx_plane_histogram(const bbx_grid_desc *grid, int n, const void *pos, int dtype, long long *plane_counts){
    if(!grid || !plane_counts || (n > 0 && !pos) || (dtype != BBX_F32 && dtype != BBX_F64)) return set_error(BBX_ERR_INVALID, "bad arguments");
    for(int z = 0; z < grid->n[2]; z++) plane_counts[z] = 0;
    for(int i = 0; i < n; i++){
        double p = dtype == BBX_F64 ? ((const double *)pos)[3 * (size_t)i + 2] : (double)((const float *)pos)[3 * (size_t)i + 2];
        double eps = 0.0;
        if(fabs(p - grid->min[2]) < 1e-8) eps = (double)0.0001f;
        else if(fabs(p - grid->max[2]) < 1e-8) eps = -(double)0.0001f;
        int u = (int)floor((p + eps - grid->min[2]) / grid->cell_len[2]);
        u = u < 0 ? 0 : (u >= grid->n[2] ? grid->n[2] - 1 : u);
        plane_counts[u]++;
    }
    return BBX_OK;
}
