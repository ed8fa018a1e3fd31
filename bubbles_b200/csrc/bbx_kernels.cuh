// bbx kernels: neighbour-grid build and the PCISPH / SPH neighbour sums, hand-written for sm_100a.
//
// Data layout in HBM (all arrays in *sorted order*: by cell id, then by the reference's chain rank):
//   pos[i]   float4  x, y, z, (unused)
//   vel[i]   float4  vx, vy, vz, rho_i          (density rides in .w so that the viscosity sum gathers
//                                                x_j, v_j, rho_j with two 16 B loads)
//   pid[i]   int     original particle id
//   cell[i]  int     cell id of slot i          (= the particle's "old cell" at the next grid update)
//   cell_start[c]    first slot of cell c (total+1 entries)
//   nbr      u16 list entries (run << 12 | offset in run), chunk-transposed: the 8 entries k..k+7 of the
//            32 particles of a warp are stored as 32 consecutive uint4 -> fully coalesced 512 B loads
//   nbr_cnt  int     Bucket::Count()
//   force, pred, posq float4; pressure, rho_pred, rho_err float
//
// The 27-cell stencil of a particle is covered by 9 "runs": for each (dy, dz) the three cells
// (cx-1..cx+1, cy+dy, cz+dz) are contiguous in memory because x is the fastest-varying cell index
// (LinearIndex, src/core/grid.h:182-193).  A list entry addresses a neighbour as (run, offset).
#pragma once
#include "bbx_device.cuh"

#define BBX_BS 128  // threads per CTA of the particle kernels

// ------------------------------------------------------------------------- A: neighbour-grid build

// A1: cell hash per particle + histogram + jump detection.
__global__ void __launch_bounds__(256) k_hash_count(int n, const float4 *__restrict__ pos, const int *__restrict__ oldcell,
                                                    int *__restrict__ newcell, int *__restrict__ count,
                                                    DevGrid g, DevState *st, int have_old)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    float4 p = pos[i];
    int ux, uy, uz;
    int c = bbx_hash(g, p.x, p.y, p.z, &ux, &uy, &uz);
    if(c < 0){
        // outside the domain: the reference would index out of bounds here (AssertA compiled out);
        // clamp to the nearest cell and raise the sticky error flag
        ux = min(max(ux, 0), g.n[0] - 1); uy = min(max(uy, 0), g.n[1] - 1); uz = min(max(uz, 0), g.n[2] - 1);
        c = ux + uy * g.n[0] + uz * g.plane;
        st->error = BBX_ERR_OUT_OF_DOMAIN;
    }
    newcell[i] = c;
    atomicAdd(&count[c], 1);
    if(have_old){
        int oc = oldcell[i];
        int oz = oc / g.plane; int rem = oc - oz * g.plane; int oy = rem / g.n[0]; int ox = rem - oy * g.n[0];
        if(abs(ox - ux) > 1 || abs(oy - uy) > 1 || abs(oz - uz) > 1){ st->jump_flag = 1; atomicAdd(&st->lost, 1); }
    }
}

// A2: exclusive scan of the per-cell counts (three small kernels: tile sums, scan of sums, tile scan)
#define SCAN_TILE 2048
__global__ void __launch_bounds__(256) k_scan_tile_sums(const int *__restrict__ count, int total, int *__restrict__ sums){
    __shared__ int ws[8];
    int base = blockIdx.x * SCAN_TILE;
    int s = 0;
#pragma unroll
    for(int k = 0; k < SCAN_TILE / 256; k++){
        int idx = base + k * 256 + threadIdx.x;
        if(idx < total) s += count[idx];
    }
    for(int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if(threadIdx.x == 0){ int t = 0; for(int k = 0; k < 8; k++) t += ws[k]; sums[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024) k_scan_sums(int *sums, int nb){
    // single CTA, sequential over chunks of 1024 with a warp-shuffle scan inside
    __shared__ int ws[32];
    __shared__ int carry;
    if(threadIdx.x == 0) carry = 0;
    __syncthreads();
    for(int base = 0; base < nb; base += 1024){
        int idx = base + threadIdx.x;
        int v = idx < nb ? sums[idx] : 0;
        int x = v;
        for(int o = 1; o < 32; o <<= 1){ int y = __shfl_up_sync(0xffffffffu, x, o); if((threadIdx.x & 31) >= o) x += y; }
        if((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
        __syncthreads();
        if(threadIdx.x < 32){
            int w = ws[threadIdx.x];
            for(int o = 1; o < 32; o <<= 1){ int y = __shfl_up_sync(0xffffffffu, w, o); if(threadIdx.x >= o) w += y; }
            ws[threadIdx.x] = w;
        }
        __syncthreads();
        int prefix = carry + (threadIdx.x >= 32 ? ws[(threadIdx.x >> 5) - 1] : 0) + x - v;
        if(idx < nb) sums[idx] = prefix;
        __syncthreads();
        if(threadIdx.x == 1023) carry = prefix + v;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) k_scan_tiles(const int *__restrict__ count, int total, const int *__restrict__ sums,
                                                    int *__restrict__ start, int n_total)
{
    __shared__ int ws[8];
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * (SCAN_TILE / 256);
    int v[SCAN_TILE / 256]; int s = 0;
#pragma unroll
    for(int k = 0; k < SCAN_TILE / 256; k++){ int idx = base + k; v[k] = idx < total ? count[idx] : 0; s += v[k]; }
    int x = s;
    for(int o = 1; o < 32; o <<= 1){ int y = __shfl_up_sync(0xffffffffu, x, o); if((threadIdx.x & 31) >= o) x += y; }
    if((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
    __syncthreads();
    int woff = 0;
    for(int k = 0; k < (int)(threadIdx.x >> 5); k++) woff += ws[k];
    int run = sums[blockIdx.x] + woff + x - s;
#pragma unroll
    for(int k = 0; k < SCAN_TILE / 256; k++){ int idx = base + k; if(idx < total){ start[idx] = run; run += v[k]; } }
    if(blockIdx.x == 0 && threadIdx.x == 0) start[total] = n_total;
}

// A3 (incremental path): one warp per *new* cell c walks the <= 27 old segments in the reference's
// neighbour order (y outer, x middle, z inner: Grid::GetNeighborListFor, grid.h:555-593) and appends, in
// old chain order, the particles whose new cell is c (Grid::DistributeToCellOpt, grid.h:449-494) --
// ballot/popc compaction, no atomics, so the order is deterministic and equal to the reference's.
// The matching lanes move the particle payload straight into the new sorted arrays.
__global__ void __launch_bounds__(256) k_fill_incremental(DevGrid g, const DevState *st,
        const int *__restrict__ start_old, const int *__restrict__ start_new, const int *__restrict__ newcell,
        const float4 *__restrict__ pos_old, const float4 *__restrict__ vel_old, const int *__restrict__ pid_old,
        float4 *__restrict__ pos_new, float4 *__restrict__ vel_new, int *__restrict__ pid_new, int *__restrict__ cell_new)
{
    if(st->rebuild_flag | st->jump_flag) return; // full rebuild path takes over
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if(warp >= g.total) return;
    int c = warp;
    int dst = start_new[c];
    int want = start_new[c + 1] - dst;
    if(want == 0) return;
    int cz = c / g.plane; int rem = c - cz * g.plane; int cy = rem / g.n[0]; int cx = rem - cy * g.n[0];
    // lane k < 27 owns neighbour k = (dy, dx, dz) in reference order
    int seg_s = 0, seg_len = 0;
    if(lane < 27){
        int dy = lane / 9 - 1, dx = (lane / 3) % 3 - 1, dz = lane % 3 - 1;
        int x = cx + dx, y = cy + dy, z = cz + dz;
        if(x >= 0 && x < g.n[0] && y >= 0 && y < g.n[1] && z >= 0 && z < g.n[2]){
            int nb = x + y * g.n[0] + z * g.plane;
            seg_s = start_old[nb];
            seg_len = start_old[nb + 1] - seg_s;
        }
    }
    int found = 0;
    for(int k = 0; k < 27 && found < want; k++){
        int s = __shfl_sync(0xffffffffu, seg_s, k);
        int len = __shfl_sync(0xffffffffu, seg_len, k);
        for(int b = 0; b < len; b += 32){
            int j = s + b + lane;
            bool m = (b + lane < len) && (newcell[j] == c);
            unsigned bal = __ballot_sync(0xffffffffu, m);
            if(m){
                int d = dst + found + __popc(bal & lanemask_lt());
                pos_new[d] = pos_old[j];
                vel_new[d] = vel_old[j];
                pid_new[d] = pid_old[j];
                cell_new[d] = c;
            }
            found += __popc(bal);
        }
    }
}

// A3' (full rebuild path, rare: Setup and the big-move rule): chains in ascending particle id
// (Grid::DistributeByParticle, grid.h:390-407).  scatter with atomics -> per-cell sort by id -> gather.
__global__ void __launch_bounds__(256) k_full_scatter(int n, const DevState *st, int force, const int *__restrict__ newcell,
        const int *__restrict__ start_new, int *__restrict__ cursor, int *__restrict__ perm)
{
    if(!force && !(st->rebuild_flag | st->jump_flag)) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    int c = newcell[i];
    int k = atomicAdd(&cursor[c], 1);
    perm[start_new[c] + k] = i;
}
__global__ void __launch_bounds__(256) k_full_sort_cells(int total, const DevState *st, int force, const int *__restrict__ start_new,
        const int *__restrict__ pid_old, int *__restrict__ perm)
{
    if(!force && !(st->rebuild_flag | st->jump_flag)) return;
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if(c >= total) return;
    int s = start_new[c], e = start_new[c + 1];
    for(int a = s + 1; a < e; a++){ // insertion sort by original id (segments are a dozen long)
        int pa = perm[a]; int ka = pid_old[pa];
        int b = a - 1;
        while(b >= s && pid_old[perm[b]] > ka){ perm[b + 1] = perm[b]; b--; }
        perm[b + 1] = pa;
    }
}
__global__ void __launch_bounds__(256) k_full_gather(int n, const DevState *st, int force, const int *__restrict__ perm,
        const int *__restrict__ newcell,
        const float4 *__restrict__ pos_old, const float4 *__restrict__ vel_old, const int *__restrict__ pid_old,
        float4 *__restrict__ pos_new, float4 *__restrict__ vel_new, int *__restrict__ pid_new, int *__restrict__ cell_new)
{
    if(!force && !(st->rebuild_flag | st->jump_flag)) return;
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if(d >= n) return;
    int j = perm[d];
    pos_new[d] = pos_old[j];
    vel_new[d] = vel_old[j];
    pid_new[d] = pid_old[j];
    cell_new[d] = newcell[j];
}

// one thread: latch the flags of this grid update and clear the per-step statistics
// (data->sphpSet->ResetHigherLevel(), pcisph_solver3.cpp:47/54)
__global__ void k_step_begin(DevState *st, int force_full){
    st->full_rebuild = (force_full | st->rebuild_flag | st->jump_flag) ? 1 : 0;
    st->rebuild_flag = 0;
    st->jump_flag = 0;
    st->overflow = 0;
    st->clamped = 0;
    st->nan_count = 0;
    st->max_force_bits = 0;
    st->max_err_bits = 0;
}
__global__ void k_clear_lost(DevState *st){ st->lost = 0; }

// ------------------------------------------------------------------ run table of a particle's cell
// base[r] / end[r] of the 9 runs r = (dy+1)*3 + (dz+1) of cell c
__device__ __forceinline__ void bbx_runs(const DevGrid &g, const int *__restrict__ cell_start, int c, int *base, int *end){
    int cz = c / g.plane; int rem = c - cz * g.plane; int cy = rem / g.n[0]; int cx = rem - cy * g.n[0];
    int xlo = max(cx - 1, 0), xhi = min(cx + 1, g.n[0] - 1);
#pragma unroll
    for(int r = 0; r < 9; r++){
        int y = cy + r / 3 - 1, z = cz + r % 3 - 1;
        if(y >= 0 && y < g.n[1] && z >= 0 && z < g.n[2]){
            int row = y * g.n[0] + z * g.plane;
            base[r] = cell_start[row + xlo];
            end[r] = cell_start[row + xhi + 1];
        }else{ base[r] = 0; end[r] = 0; }
    }
}

__device__ __forceinline__ void bbx_store_entry(unsigned short *__restrict__ nbr, int i, int k, unsigned e){
    // chunk-transposed address of entry k of particle i
    size_t warp = (size_t)(i >> 5); int lane = i & 31;
    nbr[((warp * BBX_NBR_CHUNKS + (k >> 3)) * 32 + lane) * 8 + (k & 7)] = (unsigned short)e;
}

// ---------------------------------------------------------- B: neighbour lists + density (sweep 1)
// One thread per particle.  Walks the 9 runs, tests every candidate with the reference's IsWithinStd
// predicate (bit-exact, see bbx_accept), stores the accepted ones as compact list entries and
// accumulates rho_i = m * sum W_std (ComputeDensityFor, sph_equations3.cpp:25-58).  The stored list is
// the reference's per-particle Bucket (grid.h:422-447) up to ordering; when a particle has more than
// 100 neighbours the slow path re-walks the 27 cells in the reference's order and keeps the first 100
// exactly like Bucket::Insert (particle.h:44-50).
template<int SPH_EOS>
__global__ void __launch_bounds__(BBX_BS) k_build_density(StepParams P, DevGrid g, DevState *st,
        const float4 *__restrict__ pos, float4 *__restrict__ vel, const int *__restrict__ cell,
        const int *__restrict__ cell_start, unsigned short *__restrict__ nbr, int *__restrict__ nbr_cnt,
        float *__restrict__ pressure, float4 *__restrict__ posq)
{
    int i = blockIdx.x * BBX_BS + threadIdx.x;
    if(i >= P.n) return;
    float4 pi = pos[i];
    int c = cell[i];
    int base[9], end[9];
    bbx_runs(g, cell_start, c, base, end);
    int cnt = 0; bool over = false; float sum = 0.f;
#pragma unroll 1
    for(int r = 0; r < 9 && !over; r++){
        int b = base[r], e = end[r];
        if(e - b > BBX_MAX_RUN_LEN){ st->error = BBX_ERR_CAPACITY; e = b + BBX_MAX_RUN_LEN; }
        for(int j = b; j < e; j++){
            float4 pj = pos[j];
            float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
            float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            if(bbx_accept(P, pi, pj, d2)){
                if(cnt == BBX_MAX_NEIGHBORS){ over = true; break; }
                float x = fmaxf(0.f, 1.f - d2 * P.inv_h2);
                sum += x * x * x;
                bbx_store_entry(nbr, i, cnt, ((unsigned)r << BBX_RUN_SHIFT) | (unsigned)(j - b));
                cnt++;
            }
        }
    }
    if(over){
        // reference order: y outer, x middle, z inner; chain order inside a cell; first 100 kept
        atomicAdd(&st->overflow, 1);
        cnt = 0; sum = 0.f;
        int cz = c / g.plane; int rem = c - cz * g.plane; int cy = rem / g.n[0]; int cx = rem - cy * g.n[0];
        int xlo = max(cx - 1, 0);
        for(int dy = -1; dy <= 1; dy++) for(int dx_ = -1; dx_ <= 1; dx_++) for(int dz = -1; dz <= 1; dz++){
            int x = cx + dx_, y = cy + dy, z = cz + dz;
            if(x < 0 || x >= g.n[0] || y < 0 || y >= g.n[1] || z < 0 || z >= g.n[2]) continue;
            int nb = x + y * g.n[0] + z * g.plane;
            int r = (dy + 1) * 3 + (dz + 1);
            int rb = cell_start[xlo + y * g.n[0] + z * g.plane];
            int s = cell_start[nb], e = cell_start[nb + 1];
            for(int j = s; j < e && cnt < BBX_MAX_NEIGHBORS; j++){
                if(j - rb >= BBX_MAX_RUN_LEN) break;
                float4 pj = pos[j];
                float ddx = pi.x - pj.x, ddy = pi.y - pj.y, ddz = pi.z - pj.z;
                float d2 = fmaf(ddx, ddx, fmaf(ddy, ddy, ddz * ddz));
                if(bbx_accept(P, pi, pj, d2)){
                    float xx = fmaxf(0.f, 1.f - d2 * P.inv_h2);
                    sum += xx * xx * xx;
                    bbx_store_entry(nbr, i, cnt, ((unsigned)r << BBX_RUN_SHIFT) | (unsigned)(j - rb));
                    cnt++;
                }
            }
        }
    }
    nbr_cnt[i] = cnt;
    float rho = P.mass * P.w_std_c * sum;
    // density rides in vel.w (nobody reads vel in this kernel)
    reinterpret_cast<float *>(vel)[4 * (size_t)i + 3] = rho;
    if(SPH_EOS){
        // Tait EOS, ComputePressureValue (sph_equations3.cpp:7-18)
        float p = P.eos_scale * (powf(rho / P.rho0, P.eos_exponent) - 1.f);
        if(p < 0.f) p *= P.neg_pressure_scale;
        pressure[i] = p;
        posq[i] = make_float4(pi.x, pi.y, pi.z, p / (rho * rho));
    }
}

// ------------------------------------------------------------- list walking used by sweeps 2, 3, 4
// Each thread keeps the 9 run bases of its cell in shared memory (dynamic index by run id).
#define BBX_LIST_PROLOGUE()                                                                        \
    __shared__ int sbase[9 * BBX_BS];                                                              \
    int i = blockIdx.x * BBX_BS + threadIdx.x;                                                     \
    bool live = i < P.n;                                                                           \
    int cnt = 0;                                                                                   \
    if(live){                                                                                      \
        int base[9], end[9];                                                                       \
        bbx_runs(g, cell_start, cell[i], base, end);                                               \
        _Pragma("unroll") for(int r = 0; r < 9; r++) sbase[r * BBX_BS + threadIdx.x] = base[r];    \
        cnt = nbr_cnt[i];                                                                          \
    }                                                                                              \
    const uint4 *lp = reinterpret_cast<const uint4 *>(nbr) + ((size_t)(i >> 5) * BBX_NBR_CHUNKS) * 32 + (i & 31);

template<typename F>
__device__ __forceinline__ void bbx_for_each_neighbor(const uint4 *__restrict__ lp, int cnt, const int *sbase_col, F &&body){
    for(int c0 = 0; c0 < cnt; c0 += 8){
        uint4 ch = lp[(size_t)(c0 >> 3) * 32];
        unsigned wv[4] = {ch.x, ch.y, ch.z, ch.w};
#pragma unroll
        for(int t = 0; t < 8; t++){
            if(c0 + t < cnt){
                unsigned e = (wv[t >> 1] >> ((t & 1) * 16)) & 0xffffu;
                int j = sbase_col[(e >> BBX_RUN_SHIFT) * BBX_BS] + (int)(e & BBX_RUN_MASK);
                body(j);
            }
        }
    }
}
#define BBX_LIST_FOREACH(J, ...) bbx_for_each_neighbor(lp, cnt, sbase + threadIdx.x, [&](int J) __VA_ARGS__ );

// --------------------------------- C+D: non-pressure forces + first prediction (sweep 2)
// f_i = m g - c_drag v_i + mu m^2 sum_j (v_j - v_i) d2W_spiky(d) / rho_j   (ComputeNonPressureForceFor,
// sph_equations3.cpp:80-110), then x* = x + dt (v + dt/m f), collide (restitution 0)
// (PredictVelocityAndPositionFor with is_first, pcisph_equations3.cpp:3-28).
__global__ void __launch_bounds__(BBX_BS) k_force_np_predict(StepParams P, DevGrid g, const DevColliderSet *__restrict__ cs,
        const float4 *__restrict__ pos, const float4 *__restrict__ vel, const int *__restrict__ cell,
        const int *__restrict__ cell_start, const unsigned short *__restrict__ nbr, const int *__restrict__ nbr_cnt,
        float4 *__restrict__ force, float4 *__restrict__ pred)
{
    BBX_LIST_PROLOGUE();
    if(!live) return;
    float4 pi = pos[i]; float4 vi = vel[i];
    float ax = 0.f, ay = 0.f, az = 0.f;
    BBX_LIST_FOREACH(j, {
        float4 pj = pos[j]; float4 vj = vel[j];
        float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        float d = sqrtf(d2);
        float x = fmaxf(0.f, 1.f - d * P.inv_h);
        float w = __fdividef(x, vj.w);
        ax = fmaf(vj.x - vi.x, w, ax); ay = fmaf(vj.y - vi.y, w, ay); az = fmaf(vj.z - vi.z, w, az);
    })
    float s = P.viscosity * P.mass2 * P.d2w_spiky_c;
    float fx = P.mass * P.gx - P.drag * vi.x + s * ax;
    float fy = P.mass * P.gy - P.drag * vi.y + s * ay;
    float fz = P.mass * P.gz - P.drag * vi.z + s * az;
    force[i] = make_float4(fx, fy, fz, 0.f);
    float k = P.dt * P.inv_mass;
    float tvx = vi.x + k * fx, tvy = vi.y + k * fy, tvz = vi.z + k * fz;
    float tpx = pi.x + P.dt * tvx, tpy = pi.y + P.dt * tvy, tpz = pi.z + P.dt * tvz;
    bbx_resolve_collision(*cs, (double)P.radius, 0.0, &tpx, &tpy, &tpz, &tvx, &tvy, &tvz);
    pred[i] = make_float4(tpx, tpy, tpz, 0.f);
}

// later iterations of the predict-correct loop ("correct" mode): x* from f_np + f_p, no neighbour sum
__global__ void __launch_bounds__(256) k_predict_again(StepParams P, const DevColliderSet *__restrict__ cs,
        const float4 *__restrict__ pos, const float4 *__restrict__ vel, const float4 *__restrict__ force,
        const float4 *__restrict__ force_p, float4 *__restrict__ pred)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= P.n) return;
    float4 pi = pos[i], vi = vel[i], f = force[i], fp = force_p[i];
    float k = P.dt * P.inv_mass;
    float tvx = vi.x + k * (f.x + fp.x), tvy = vi.y + k * (f.y + fp.y), tvz = vi.z + k * (f.z + fp.z);
    float tpx = pi.x + P.dt * tvx, tpy = pi.y + P.dt * tvy, tpz = pi.z + P.dt * tvz;
    bbx_resolve_collision(*cs, (double)P.radius, 0.0, &tpx, &tpy, &tpz, &tvx, &tvy, &tvz);
    pred[i] = make_float4(tpx, tpy, tpz, 0.f);
}

// ----------------------------------------------------------- E: predicted density -> pressure (sweep 3)
// rho*_i = m sum_j W_std(|x*_i - x*_j|) over the same list; p += delta (rho* - rho0), negative increments
// scaled by negativePressureScale (PredictPressureFor, pcisph_equations3.cpp:60-90).
// Writes posq = (x_i, p_i / rho*_i^2) for the pressure-force sweep.
__global__ void __launch_bounds__(BBX_BS) k_pressure(StepParams P, DevGrid g, DevState *st, int first,
        const float4 *__restrict__ pos, const float4 *__restrict__ pred, const int *__restrict__ cell,
        const int *__restrict__ cell_start, const unsigned short *__restrict__ nbr, const int *__restrict__ nbr_cnt,
        float *__restrict__ pressure, float *__restrict__ rho_pred, float *__restrict__ rho_err, float4 *__restrict__ posq)
{
    BBX_LIST_PROLOGUE();
    if(!live) return;
    float4 pi = pred[i];
    float sum = 0.f;
    BBX_LIST_FOREACH(j, {
        float4 pj = pred[j];
        float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        float x = fmaxf(0.f, 1.f - d2 * P.inv_h2);
        sum = fmaf(x * x, x, sum);
    })
    float rho = P.mass * P.w_std_c * sum;
    float err = rho - P.rho0;
    float dp = P.delta * err;
    if(dp < 0.f){ dp *= P.neg_pressure_scale; err *= P.neg_pressure_scale; }
    float p = (first ? 0.f : pressure[i]) + dp;
    pressure[i] = p;
    rho_pred[i] = rho;
    rho_err[i] = err;
    float4 x0 = pos[i];
    float rho2 = rho * rho;
    // the reference skips a neighbour whose rho*^2 is ~0 (pcisph_equations3.cpp:137): NaN marks it
    posq[i] = make_float4(x0.x, x0.y, x0.z, (rho2 < 1e-8f) ? __int_as_float(0x7fc00000) : p / rho2);
    atomicMax(&st->max_err_bits, __float_as_uint(fabsf(err)));
}

// -------------------------------------- F+G: pressure force (+ accumulate, integrate, collide) (sweep 4)
// f_p,i = - m^2 sum_{j != i} (p_i/rho*_i^2 + p_j/rho*_j^2) gradW_spiky  with current positions
// (PredictPressureForceFor, pcisph_equations3.cpp:114-155); INTEGRATE: f += f_p, v += dt f/m, x += dt v,
// collide (restitution 0.6), domain clamp, big-move flag (AccumulateForcesFor + TimeIntegrationFor,
// pcisph_equations3.cpp:179-197, sph_equations3.cpp:283-339).  Positions are updated in place: neighbours
// are read from posq, never from pos, so there is no read/write race.
template<int INTEGRATE>
__global__ void __launch_bounds__(BBX_BS) k_pressure_force(StepParams P, DevGrid g, DevState *st, const DevColliderSet *__restrict__ cs,
        float4 *__restrict__ pos, float4 *__restrict__ vel, const float4 *__restrict__ posq, const int *__restrict__ cell,
        const int *__restrict__ cell_start, const unsigned short *__restrict__ nbr, const int *__restrict__ nbr_cnt,
        float4 *__restrict__ force, float4 *__restrict__ force_p)
{
    BBX_LIST_PROLOGUE();
    if(!live) return;
    float4 pi = posq[i];
    float tx = 0.f, ty = 0.f, tz = 0.f;
    float qi = pi.w;
    BBX_LIST_FOREACH(j, {
        float4 pj = posq[j];
        float dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
        float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        float d = sqrtf(d2);
        float x = fmaxf(0.f, 1.f - d * P.inv_h);
        // j == i, coincident points and NaN-marked neighbours contribute nothing
        bool ok = (j != i) && (d > 1e-8f) && (pj.w == pj.w);
        float w = ok ? __fdividef((qi + pj.w) * x * x, d) : 0.f;
        tx = fmaf(dx, w, tx); ty = fmaf(dy, w, ty); tz = fmaf(dz, w, tz);
    })
    float s = -P.mass2 * P.dw_spiky_c;
    float fpx = s * tx, fpy = s * ty, fpz = s * tz;
    force_p[i] = make_float4(fpx, fpy, fpz, 0.f);
    if(INTEGRATE){
        float4 f = force[i]; float4 v = vel[i];
        float fx = f.x + fpx, fy = f.y + fpy, fz = f.z + fpz;
        force[i] = make_float4(fx, fy, fz, 0.f);
        float vx = v.x + P.dt * (fx * P.inv_mass), vy = v.y + P.dt * (fy * P.inv_mass), vz = v.z + P.dt * (fz * P.inv_mass);
        float px = pi.x + P.dt * vx, py = pi.y + P.dt * vy, pz = pi.z + P.dt * vz;
        bbx_resolve_collision(*cs, (double)P.radius, (double)P.restitution, &px, &py, &pz, &vx, &vy, &vz);
        // domain clamp (sph_equations3.cpp:317-326)
        if(!inside_bounds(v3(px, py, pz), v3(g.min[0], g.min[1], g.min[2]), v3(g.max[0], g.max[1], g.max[2]))){
            double r = (double)P.radius;
            px = (float)clampd(px, g.min[0] + r, g.max[0] - r);
            py = (float)clampd(py, g.min[1] + r, g.max[1] - r);
            pz = (float)clampd(pz, g.min[2] + r, g.max[2] - r);
            atomicAdd(&st->clamped, 1);
        }
        float mx = px - pi.x, my = py - pi.y, mz = pz - pi.z;
        if(sqrtf(mx * mx + my * my + mz * mz) >= P.min_cell_len09) st->rebuild_flag = 1;
        if(!(isfinite(px) && isfinite(py) && isfinite(pz))) atomicAdd(&st->nan_count, 1);
        pos[i] = make_float4(px, py, pz, 0.f);
        vel[i] = make_float4(vx, vy, vz, v.w);
        atomicMax(&st->max_force_bits, __float_as_uint(sqrtf(fx * fx + fy * fy + fz * fz)));
    }
}

// integrate alone ("correct" mode after the loop, and the SPH step)
__global__ void __launch_bounds__(256) k_integrate(StepParams P, DevGrid g, DevState *st, const DevColliderSet *__restrict__ cs,
        float4 *__restrict__ pos, float4 *__restrict__ vel, float4 *__restrict__ force, const float4 *__restrict__ force_p)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= P.n) return;
    float4 pi = pos[i], v = vel[i], f = force[i];
    float fx = f.x, fy = f.y, fz = f.z;
    if(force_p){ float4 fp = force_p[i]; fx += fp.x; fy += fp.y; fz += fp.z; force[i] = make_float4(fx, fy, fz, 0.f); }
    float vx = v.x + P.dt * (fx * P.inv_mass), vy = v.y + P.dt * (fy * P.inv_mass), vz = v.z + P.dt * (fz * P.inv_mass);
    float px = pi.x + P.dt * vx, py = pi.y + P.dt * vy, pz = pi.z + P.dt * vz;
    bbx_resolve_collision(*cs, (double)P.radius, (double)P.restitution, &px, &py, &pz, &vx, &vy, &vz);
    if(!inside_bounds(v3(px, py, pz), v3(g.min[0], g.min[1], g.min[2]), v3(g.max[0], g.max[1], g.max[2]))){
        double r = (double)P.radius;
        px = (float)clampd(px, g.min[0] + r, g.max[0] - r);
        py = (float)clampd(py, g.min[1] + r, g.max[1] - r);
        pz = (float)clampd(pz, g.min[2] + r, g.max[2] - r);
        atomicAdd(&st->clamped, 1);
    }
    float mx = px - pi.x, my = py - pi.y, mz = pz - pi.z;
    if(sqrtf(mx * mx + my * my + mz * mz) >= P.min_cell_len09) st->rebuild_flag = 1;
    if(!(isfinite(px) && isfinite(py) && isfinite(pz))) atomicAdd(&st->nan_count, 1);
    pos[i] = make_float4(px, py, pz, 0.f);
    vel[i] = make_float4(vx, vy, vz, v.w);
    atomicMax(&st->max_force_bits, __float_as_uint(sqrtf(fx * fx + fy * fy + fz * fz)));
}

// ------------------------------------------------------------------ SPH (non-PCI) force sweep
// ComputeAllForcesFor (sph_equations3.cpp:184-272) with Jacobi semantics: gravity + drag + viscosity
// (only inside the spiky support and j != i) + pressure force from rho, p of this sub-step.
__global__ void __launch_bounds__(BBX_BS) k_sph_forces(StepParams P, DevGrid g,
        const float4 *__restrict__ posq, const float4 *__restrict__ vel, const int *__restrict__ cell,
        const int *__restrict__ cell_start, const unsigned short *__restrict__ nbr, const int *__restrict__ nbr_cnt,
        float4 *__restrict__ force)
{
    BBX_LIST_PROLOGUE();
    if(!live) return;
    float4 pi = posq[i]; float4 vi = vel[i];
    float tx = 0.f, ty = 0.f, tz = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
    float qi = pi.w;
    BBX_LIST_FOREACH(j, {
        float4 pj = posq[j]; float4 vj = vel[j];
        float dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
        float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        float d = sqrtf(d2);
        float x = (j != i) ? fmaxf(0.f, 1.f - d * P.inv_h) : 0.f;
        float w = (d > 1e-8f) ? __fdividef((qi + pj.w) * x * x, d) : 0.f;
        tx = fmaf(dx, w, tx); ty = fmaf(dy, w, ty); tz = fmaf(dz, w, tz);
        float wv = __fdividef(x, vj.w);
        ax = fmaf(vj.x - vi.x, wv, ax); ay = fmaf(vj.y - vi.y, wv, ay); az = fmaf(vj.z - vi.z, wv, az);
    })
    float sp = -P.mass2 * P.dw_spiky_c;
    float sv = P.viscosity * P.mass2 * P.d2w_spiky_c;
    force[i] = make_float4(P.mass * P.gx - P.drag * vi.x + sv * ax + sp * tx,
                           P.mass * P.gy - P.drag * vi.y + sv * ay + sp * ty,
                           P.mass * P.gz - P.drag * vi.z + sv * az + sp * tz, 0.f);
}

// ------------------------------------------------------------------ pseudo-viscosity (cold branch)
// ComputePseudoViscosity{Aggregation,Interpolation}KernelFor (sph_equations3.cpp:341-382): only runs
// when pseudoViscosity * dt > 0.1 (never with the default dt).
__global__ void __launch_bounds__(BBX_BS) k_pseudo_aggregate(StepParams P, DevGrid g,
        const float4 *__restrict__ pos, const float4 *__restrict__ vel, const int *__restrict__ cell,
        const int *__restrict__ cell_start, const unsigned short *__restrict__ nbr, const int *__restrict__ nbr_cnt,
        float4 *__restrict__ smoothed)
{
    BBX_LIST_PROLOGUE();
    if(!live) return;
    float4 pi = pos[i];
    float sx = 0.f, sy = 0.f, sz = 0.f, ws = 0.f;
    BBX_LIST_FOREACH(j, {
        float4 pj = pos[j]; float4 vj = vel[j];
        float dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
        float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
        float x = fmaxf(0.f, 1.f - d * P.inv_h);
        float w = P.mass / vj.w * (P.w_spiky_c * x * x * x);
        ws += w; sx = fmaf(w, vj.x, sx); sy = fmaf(w, vj.y, sy); sz = fmaf(w, vj.z, sz);
    })
    if(ws > 0.f){ float inv = 1.f / ws; sx *= inv; sy *= inv; sz *= inv; }
    smoothed[i] = make_float4(sx, sy, sz, 0.f);
}
__global__ void __launch_bounds__(256) k_pseudo_interpolate(StepParams P, float4 *__restrict__ vel, const float4 *__restrict__ smoothed){
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= P.n) return;
    float4 v = vel[i], s = smoothed[i];
    float t = P.pseudo_factor;
    vel[i] = make_float4((1.f - t) * v.x + t * s.x, (1.f - t) * v.y + t * s.y, (1.f - t) * v.z + t * s.z, v.w);
}

// ------------------------------------------------------------------ upload / download / export
__global__ void __launch_bounds__(256) k_upload(int n, int first_id, const void *__restrict__ pos, const void *__restrict__ vel, int is_f64,
                                                float4 *__restrict__ dpos, float4 *__restrict__ dvel, int *__restrict__ pid)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    float p[3], v[3];
    for(int k = 0; k < 3; k++){
        if(is_f64){ p[k] = (float)((const double *)pos)[3 * (size_t)i + k]; v[k] = (float)((const double *)vel)[3 * (size_t)i + k]; }
        else{ p[k] = ((const float *)pos)[3 * (size_t)i + k]; v[k] = ((const float *)vel)[3 * (size_t)i + k]; }
    }
    dpos[i] = make_float4(p[0], p[1], p[2], 0.f);
    dvel[i] = make_float4(v[0], v[1], v[2], 0.f);
    pid[i] = first_id + i;
}
// overwrite pos/vel of existing particles: slot i holds particle pid[i]
__global__ void __launch_bounds__(256) k_overwrite(int n, const int *__restrict__ pid, const void *__restrict__ pos, const void *__restrict__ vel,
                                                   int is_f64, float4 *__restrict__ dpos, float4 *__restrict__ dvel)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    size_t id = (size_t)pid[i];
    float p[3], v[3];
    for(int k = 0; k < 3; k++){
        if(is_f64){ p[k] = (float)((const double *)pos)[3 * id + k]; v[k] = (float)((const double *)vel)[3 * id + k]; }
        else{ p[k] = ((const float *)pos)[3 * id + k]; v[k] = ((const float *)vel)[3 * id + k]; }
    }
    dpos[i] = make_float4(p[0], p[1], p[2], dpos[i].w);
    dvel[i] = make_float4(v[0], v[1], v[2], dvel[i].w);
}
// scatter a sorted-order field back to original-id order. comp: 3 = xyz of a float4 array, 1 = .w of a
// float4 array (src4) or a plain float array (src1)
__global__ void __launch_bounds__(256) k_download(int n, const int *__restrict__ pid, const float4 *__restrict__ src4,
                                                  const float *__restrict__ src1, const int *__restrict__ srci, int comps, int is_f64, void *__restrict__ dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    size_t id = (size_t)pid[i];
    if(srci){ ((int *)dst)[id] = srci[i]; return; }
    if(comps == 3){
        float4 v = src4[i];
        if(is_f64){ double *d = (double *)dst + 3 * id; d[0] = v.x; d[1] = v.y; d[2] = v.z; }
        else{ float *d = (float *)dst + 3 * id; d[0] = v.x; d[1] = v.y; d[2] = v.z; }
    }else{
        float v = src4 ? src4[i].w : src1[i];
        if(is_f64) ((double *)dst)[id] = v; else ((float *)dst)[id] = v;
    }
}
__global__ void __launch_bounds__(256) k_export_cells(int total, const int *__restrict__ cell_start, int *__restrict__ cell_count){
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if(c < total) cell_count[c] = cell_start[c + 1] - cell_start[c];
}
// Decode the *stored* compact lists into the reference's bucket order (neighbour cells y/x/z, chain
// order inside a cell) with original ids: what Bucket::pids holds after UpdateParticlesBuckets.
__global__ void __launch_bounds__(BBX_BS) k_export_neighbors(int n, DevGrid g, const int *__restrict__ pid, const int *__restrict__ cell,
        const int *__restrict__ cell_start, const unsigned short *__restrict__ nbr, const int *__restrict__ nbr_cnt,
        int *__restrict__ counts, int *__restrict__ ids)
{
    int i = blockIdx.x * BBX_BS + threadIdx.x;
    if(i >= n) return;
    int c = cell[i];
    int base[9], end[9];
    bbx_runs(g, cell_start, c, base, end);
    int cz = c / g.plane; int rem = c - cz * g.plane; int cy = rem / g.n[0]; int cx = rem - cy * g.n[0];
    int cnt = nbr_cnt[i];
    size_t id = (size_t)pid[i];
    int *out = ids + id * BBX_MAX_NEIGHBORS;
    // key = (reference rank of the neighbour cell) << 20 | slot: sort ascending (insertion, <= 100 items)
    unsigned long long keys[BBX_MAX_NEIGHBORS];
    for(int k = 0; k < cnt; k++){
        size_t warp = (size_t)(i >> 5); int lane = i & 31;
        unsigned e = nbr[((warp * BBX_NBR_CHUNKS + (k >> 3)) * 32 + lane) * 8 + (k & 7)];
        int r = e >> BBX_RUN_SHIFT; int j = base[r] + (int)(e & BBX_RUN_MASK);
        int cj = cell[j];
        int jx = cj % g.n[0];
        int dy = r / 3 - 1, dz = r % 3 - 1, dx = jx - cx;
        unsigned rank = (unsigned)((dy + 1) * 9 + (dx + 1) * 3 + (dz + 1));
        unsigned long long key = ((unsigned long long)rank << 40) | (unsigned long long)(unsigned)j;
        int b = k - 1;
        while(b >= 0 && keys[b] > key){ keys[b + 1] = keys[b]; b--; }
        keys[b + 1] = key;
    }
    for(int k = 0; k < BBX_MAX_NEIGHBORS; k++) out[k] = k < cnt ? pid[(int)(keys[k] & 0xffffffffffULL)] : -1;
    counts[id] = cnt;
}
// inject chains: slot d takes particle order[d] (original id) -> need id -> old slot map
__global__ void __launch_bounds__(256) k_slot_of_id(int n, const int *__restrict__ pid, int *__restrict__ slot_of){
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) slot_of[pid[i]] = i;
}
__global__ void __launch_bounds__(256) k_inject_gather(int n, const int *__restrict__ order, const int *__restrict__ slot_of,
        const int *__restrict__ cell_of_slot_new,
        const float4 *__restrict__ pos_old, const float4 *__restrict__ vel_old,
        float4 *__restrict__ pos_new, float4 *__restrict__ vel_new, int *__restrict__ pid_new, int *__restrict__ cell_new)
{
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if(d >= n) return;
    int id = order[d]; int j = slot_of[id];
    pos_new[d] = pos_old[j]; vel_new[d] = vel_old[j]; pid_new[d] = id; cell_new[d] = cell_of_slot_new[d];
}
