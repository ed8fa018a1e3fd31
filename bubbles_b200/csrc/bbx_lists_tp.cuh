// bbx list build, v11 (phase B): per-particle neighbour lists + density, ONE THREAD PER PARTICLE, one warp per TILE of 32
// consecutive slots, the candidates of the tile staged once in shared memory.
//
// Why this shape (profiles/r02_notes.md): the warp-per-cell kernel (bbx_lists.cuh, v7) spends 16 instructions per
// (32 candidates, 1 particle) in its pair loop -- and as much again staging ~313 candidates per cell, of which 2/3 are
// re-staged by the next cell of the x-row, and copying 4 list rows out per group.  Here
//   * slots are sorted by cell and x is the fastest cell index, so the 32 particles of a tile sit in ~4 consecutive cells of
//     one x-row and, for each of the 9 (dy, dz) offsets, everything they can see is ONE contiguous slot range: 9 bulk copies
//     (cp.async.bulk + mbarrier, UBLKCP in SASS) bring the ~570 raw positions in -- ~18 candidates staged per particle
//     instead of ~313 per 10, and no per-candidate copy instruction at all;
//   * every lane then walks ITS OWN 3-cell window of each run in lock step with the others (lanes of one cell read the same
//     address: a broadcast), so the list order is the walk order and the 100 cap and the FP64 band check are thread-local;
//   * the tile's 32 list rows are exactly one chunk-transposed block of `nbr`: the copy-out is coalesced 512-byte stores.
// A tile that spans two x-rows runs one pass per row; a group of lanes whose union exceeds the stage is halved until it
// fits; a single particle whose own 27 cells exceed it, and a particle that runs into the 100 cap, is redone by
// bbx_list_walk_exact straight from global memory in the reference's traversal order.
//
// Arithmetic: d^2 straight from the raw FP32 positions in one fixed operation order, so every value (acceptance, density
// weight) depends on the pair alone and not on the tile it was computed in: a slab engine, whose slots are numbered
// differently, reproduces the single-domain engine bit for bit.  d^2 < thr_lo: certainly inside; thr_lo <= d^2 <= thr_hi (the
// guard band around h^2 - 1e-8): the reference's FP64 predicate decides on the spot (IsWithinStd, kernel.cpp:229-234).
// Reference: Grid::DistributeParticleBucket grid.h:422-447, Bucket::Insert particle.h:44-50 (cap 100),
// ComputeDensityFor + ComputePressureValue sph_equations3.cpp:7-58.
#pragma once
#include "bbx_device.cuh"

#define BBX_TP_WARPS 4            // warps (tiles in flight) per CTA
#ifndef BBX_TP_CAP
#define BBX_TP_CAP 768            // staged candidates per warp (16 B each)
#endif
#ifndef BBX_TP_UNROLL
#define BBX_TP_UNROLL 4
#endif
#define BBX_TP_ROW (96 + 8 * ((BBX_TP_UNROLL + 7) / 8) + 8 * (BBX_TP_UNROLL > 4))   // u16 entries per list row in shared memory: 100 + one block, in chunks of 8
#define BBX_TP_WARP_BYTES (BBX_TP_CAP * 16 + 32 * BBX_TP_ROW * 2 + 16)
#define BBX_TP_SMEM (BBX_TP_WARPS * BBX_TP_WARP_BYTES)
#ifndef BBX_TP_MINB
#define BBX_TP_MINB 3
#endif
#ifndef BBX_TP_RUNROLL
#define BBX_TP_RUNROLL 9          // unroll factor of the loop over the 9 runs (1: rolled, the run tables then live in local memory)
#endif
#define BBX_TP_FULL 0xffffffffu
#define BBX_TP_PRAGMA(x) _Pragma(#x)
#define BBX_TP_UNROLL_N(n) BBX_TP_PRAGMA(unroll n)

// ---- mbarrier / bulk-copy plumbing (PTX; SASS: SYNCS.*, UBLKCP)
__device__ __forceinline__ unsigned bbx_smem_u32(const void *p){ return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bbx_mbar_init(unsigned long long *b, unsigned count){
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bbx_smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void bbx_mbar_arrive_expect_tx(unsigned long long *b, unsigned bytes){
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bbx_smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bbx_mbar_wait(unsigned long long *b, unsigned parity){
    asm volatile("{\n\t.reg .pred p;\n\tBBX_WAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra BBX_DONE_%=;\n\tbra BBX_WAIT_%=;\n\tBBX_DONE_%=:\n\t}"
                 :: "r"(bbx_smem_u32(b)), "r"(parity) : "memory");
}
// global -> shared, `bytes` (multiple of 16, both addresses 16-byte aligned), completion counted on the mbarrier
__device__ __forceinline__ void bbx_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *b){
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(bbx_smem_u32(dst)), "l"(src), "r"(bytes), "r"(bbx_smem_u32(b)) : "memory");
}

// One particle straight from global memory, in the reference's traversal order (y outer, x middle, z inner; chain order
// inside a cell): keeps the first 100 accepted exactly like Bucket::Insert, the density is the sum over exactly those.
// Used for a particle that ran into the cap and for one whose 27 cells do not fit the stage.
__device__ __noinline__ void bbx_list_walk_exact(const StepParams &P, const DevGrid &g, DevState *st,
        const float4 *__restrict__ pos, const int *__restrict__ cell_start, unsigned short *row,
        int c, float4 pi, int *cnt_out, float *sum_out)
{
    int cnt = 0, total = 0; float sum = 0.f;
    int cz = c / g.plane; int rem = c - cz * g.plane; int cy = rem / g.n[0]; int cx = rem - cy * g.n[0];
    int xlo = max(cx - 1, 0);
    for(int dy = -1; dy <= 1; dy++) for(int dx_ = -1; dx_ <= 1; dx_++) for(int dz = -1; dz <= 1; dz++){
        int x = cx + dx_, y = cy + dy, z = cz + dz;
        if(x < 0 || x >= g.n[0] || y < 0 || y >= g.n[1] || z < 0 || z >= g.n[2]) continue;
        int nb = x + y * g.n[0] + z * g.plane;
        int r = (dy + 1) * 3 + (dz + 1);
        int rb = cell_start[xlo + y * g.n[0] + z * g.plane];
        int s = cell_start[nb], e = cell_start[nb + 1];
        for(int j = s; j < e; j++){
            if(j - rb >= BBX_MAX_RUN_LEN) break;
            float4 pj = pos[j];
            float ddx = pi.x - pj.x, ddy = pi.y - pj.y, ddz = pi.z - pj.z;
            float d2 = fmaf(ddx, ddx, fmaf(ddy, ddy, ddz * ddz));
            if(bbx_accept(P, pi, pj, d2)){
                total++;
                if(cnt < BBX_MAX_NEIGHBORS){
                    const float yy = P.h2 - d2;
                    sum = fmaf(yy * yy, yy, sum);
                    row[cnt] = (unsigned short)(((unsigned)r << BBX_RUN_SHIFT) | (unsigned)(j - rb));
                    cnt++;
                }
            }
        }
    }
    if(total > BBX_MAX_NEIGHBORS) atomicAdd(&st->overflow, 1);
    *cnt_out = cnt; *sum_out = sum;
}

template<int SPH_EOS>
__global__ void __launch_bounds__(BBX_TP_WARPS * 32, BBX_TP_MINB) k_lists_density_tp(const __grid_constant__ StepParams P, const __grid_constant__ DevGrid g, DevState *st,
        const int *__restrict__ cell, const float4 *__restrict__ pos, float4 *__restrict__ vel, const int *__restrict__ cell_start,
        unsigned short *__restrict__ nbr, int *__restrict__ nbr_cnt, float *__restrict__ pressure, float4 *__restrict__ posq,
        float4 *__restrict__ rec, HaloDst H)
{
    extern __shared__ __align__(16) unsigned char tp_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *wbase = tp_smem + warp * BBX_TP_WARP_BYTES;
    float4 *cand = reinterpret_cast<float4 *>(wbase);
    unsigned short *rows = reinterpret_cast<unsigned short *>(wbase + BBX_TP_CAP * 16);
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(wbase + BBX_TP_CAP * 16 + 32 * BBX_TP_ROW * 2);
    if(lane == 0){ bbx_mbar_init(mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    unsigned parity = 0;
    const int n = bbx_count(P);
    H = bbx_halo_resolve(H, st);
    const int ntiles = (n + 31) >> 5, nwarps = gridDim.x * BBX_TP_WARPS;
    unsigned short *my_row = rows + lane * BBX_TP_ROW;
    const unsigned row_addr = bbx_smem_u32(my_row);
#pragma unroll 1
    for(int t = blockIdx.x * BBX_TP_WARPS + warp; t < ntiles; t += nwarps){
        const int i = t * 32 + lane;
        const bool live = i < n;
        int c = 0; float4 pi = make_float4(0.f, 0.f, 0.f, 0.f);
        if(live){ c = cell[i]; pi = pos[i]; }
        const int rowid = c / g.n[0], cx = c - rowid * g.n[0];
        const int cz = rowid / g.n[1], cy = rowid - cz * g.n[1];
        // the lane's 9 runs: run r = (dy + 1) * 3 + (dz + 1) covers the cells (cx-1 .. cx+1, cy+dy, cz+dz), contiguous slots
        int lo[9], len[9], T = 0;
        {
            const int xlo = max(cx - 1, 0), xhi = min(cx + 1, g.n[0] - 1);
#pragma unroll
            for(int r = 0; r < 9; r++){
                const int y = cy + r / 3 - 1, z = cz + r % 3 - 1;
                lo[r] = 0; len[r] = 0;
                if(live && y >= 0 && y < g.n[1] && z >= 0 && z < g.n[2]){
                    const int rb = y * g.n[0] + z * g.plane;
                    lo[r] = cell_start[rb + xlo];
                    len[r] = cell_start[rb + xhi + 1] - lo[r];
                    if(len[r] > BBX_MAX_RUN_LEN){ st->error = BBX_ERR_CAPACITY; len[r] = BBX_MAX_RUN_LEN; }
                }
                T += len[r];
            }
        }
        {
            const int tmax = __reduce_max_sync(BBX_TP_FULL, T);
            if(lane == 0 && tmax > st->max_candidates) atomicMax(&st->max_candidates, tmax);
        }
        unsigned wa = row_addr;         // shared-memory address of the next list entry of this lane
        float acc = 0.f;
        float4 cj[BBX_TP_UNROLL];       // (a lane past the end of its window keeps stale, finite values: never accepted)
#pragma unroll
        for(int u = 0; u < BBX_TP_UNROLL; u++) cj[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        bool slow = false;
        unsigned todo = __ballot_sync(BBX_TP_FULL, live);
#pragma unroll 1
        while(todo){
            // ---- the group of this pass: the remaining lanes of the first remaining lane's x-row, halved until it fits
            int f = __ffs(todo) - 1;
            const int row_f = __shfl_sync(BBX_TP_FULL, rowid, f);
            unsigned grp = __ballot_sync(BBX_TP_FULL, live && rowid == row_f) & todo;
            int sidx[9], msrc = 0, mlen = 0, mdst = 0, U, l;
#pragma unroll 1
            for(;;){
                l = 31 - __clz(grp);
                U = 0;
#pragma unroll
                for(int r = 0; r < 9; r++){
                    const int a = __shfl_sync(BBX_TP_FULL, lo[r], f), b = __shfl_sync(BBX_TP_FULL, lo[r] + len[r], l);
                    sidx[r] = U + lo[r] - a;
                    if(lane == r){ msrc = a; mlen = b - a; mdst = U; }
                    U += b - a;
                }
                if(U <= BBX_TP_CAP || f == l) break;
                grp = ((1u << ((__popc(grp) + 1) >> 1)) - 1u) << f;   // (the lanes of a group are contiguous)
            }
            todo &= ~grp;
            if(U > BBX_TP_CAP){ if(lane == f) slow = true; continue; }   // one particle, 27 cells larger than the stage
            const bool member = (grp >> lane) & 1u;
            // ---- stage: 9 bulk copies of raw positions (one contiguous slot range per run)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy reads of the buffer first
            __syncwarp();
            if(lane == 0) bbx_mbar_arrive_expect_tx(mbar, 16u * (unsigned)U);
            __syncwarp();
            if(lane < 9 && mlen > 0) bbx_bulk_g2s(cand + mdst, pos + msrc, 16u * (unsigned)mlen, mbar);
            bbx_mbar_wait(mbar, parity);
            parity ^= 1u;
            // ---- the walk: run by run, every member lane over its own window, in lock step.  A candidate up to thr_hi is
            // appended provisionally; a lane that took one from inside the guard band redoes the run with the FP64 predicate.
            // Blocks that lie inside every member's window run without the per-step window test.
            BBX_TP_UNROLL_N(BBX_TP_RUNROLL)
            for(int r = 0; r < 9; r++){
                const int mylen = member ? len[r] : 0;
                const int L = __reduce_max_sync(BBX_TP_FULL, mylen);
                const int Lall = __reduce_min_sync(BBX_TP_FULL, member ? len[r] : 0x7fffffff) & ~(BBX_TP_UNROLL - 1);
                const float4 *cp = cand + (member ? sidx[r] : 0);   // (a lane outside the group reads along, harmlessly)
                const unsigned lim = row_addr + 2u * BBX_MAX_NEIGHBORS;
                const unsigned wa0 = wa; const float acc0 = acc;
                float dmx = 0.f;
                int k0 = 0;
#define BBX_TP_STEP(VALID)                                                                                                   \
                    {                                                                                                        \
                        const float dx = pi.x - cj[u].x, dy = pi.y - cj[u].y, dz = pi.z - cj[u].z;                           \
                        const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));                                                \
                        if((VALID) && d2 <= P.thr_hi){                                                                       \
                            const float y = P.h2 - d2;                                                                       \
                            asm volatile("st.shared.u16 [%0], %1;" :: "r"(wa), "h"((unsigned short)(((unsigned)r << BBX_RUN_SHIFT) | (unsigned)(k0 + u))) : "memory"); \
                            wa += 2u;                                                                                        \
                            acc = fmaf(y * y, y, acc);                                                                       \
                            dmx = fmaxf(dmx, d2);                                                                            \
                        }                                                                                                    \
                    }
#pragma unroll 1
                for(; k0 < Lall; k0 += BBX_TP_UNROLL){
                    // a row holds 104 entries: past 100 the lane is redone by the slow path, its row is scratch from here on
                    if(wa > lim){ slow = true; wa = row_addr; }
#pragma unroll
                    for(int u = 0; u < BBX_TP_UNROLL; u++) cj[u] = cp[k0 + u];
#pragma unroll
                    for(int u = 0; u < BBX_TP_UNROLL; u++) BBX_TP_STEP(member)
                }
#pragma unroll 1
                for(; k0 < L; k0 += BBX_TP_UNROLL){
                    if(wa > lim){ slow = true; wa = row_addr; }
#pragma unroll
                    for(int u = 0; u < BBX_TP_UNROLL; u++) if(k0 + u < mylen) cj[u] = cp[k0 + u];
#pragma unroll
                    for(int u = 0; u < BBX_TP_UNROLL; u++) BBX_TP_STEP(k0 + u < mylen)   // (mylen = 0 outside the group)
                }
#undef BBX_TP_STEP
                if(dmx >= P.thr_lo){
                    wa = wa0; acc = acc0;
#pragma unroll 1
                    for(int k = 0; k < mylen; k++){
                        if(wa > lim){ slow = true; wa = row_addr; }
                        const float4 c4 = cp[k];
                        const float dx = pi.x - c4.x, dy = pi.y - c4.y, dz = pi.z - c4.z;
                        const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        bool in = d2 < P.thr_lo;
                        if(!in && d2 <= P.thr_hi){ in = bbx_within_std_exact(pi, c4, P.h2_d); atomicAdd(&st->exact_passes, 1); }
                        if(in){
                            const float y = P.h2 - d2;
                            asm volatile("st.shared.u16 [%0], %1;" :: "r"(wa), "h"((unsigned short)(((unsigned)r << BBX_RUN_SHIFT) | (unsigned)k)) : "memory");
                            wa += 2u;
                            acc = fmaf(y * y, y, acc);
                        }
                    }
                }
            }
        }
        // ---- per particle: count, density, the records of the force sweeps
        int cnt = (int)(wa - row_addr) >> 1;
        float sum = acc;
        if(cnt > BBX_MAX_NEIGHBORS) slow = true;
        if(slow) bbx_list_walk_exact(P, g, st, pos, cell_start, my_row, c, pi, &cnt, &sum);
        if(live){
            nbr_cnt[i] = cnt;
            // W_std = c (1 - d^2 / h^2)^3 = c / h^6 (h^2 - d^2)^3: the sum runs over (h^2 - d^2)^3
            const float rho = P.mass * P.w_std_c * (P.inv_h2 * P.inv_h2 * P.inv_h2) * sum;
            // rho rides in vel.w and in the 32-byte gather record (x, y, z, rho | vx, vy, vz, -) of the viscosity
            // sweep; the grid fill wrote the record's x and v
            reinterpret_cast<float *>(vel)[4 * (size_t)i + 3] = rho;
            reinterpret_cast<float *>(rec)[8 * (size_t)i + 3] = rho;
            if(i < H.n_first || i >= H.hi_begin){
                // boundary plane of a slab: the complete record goes to the neighbour's ghost slot as well
                const float4 vv = vel[i];
                bbx_halo_store(H, 0, 2, i, 0, make_float4(pi.x, pi.y, pi.z, rho));
                bbx_halo_store(H, 0, 2, i, 1, make_float4(vv.x, vv.y, vv.z, 0.f));
            }
            if(SPH_EOS){
                // Tait EOS, ComputePressureValue (sph_equations3.cpp:7-18)
                float p = P.eos_scale * (powf(rho / P.rho0, P.eos_exponent) - 1.f);
                if(p < 0.f) p *= P.neg_pressure_scale;
                pressure[i] = p;
                const float qq = p / (rho * rho);
                posq[i] = make_float4(pi.x, pi.y, pi.z, qq);
                reinterpret_cast<float *>(rec)[8 * (size_t)i + 7] = qq; // the SPH force sweep gathers (x, rho | v, p / rho^2)
            }
        }
        // ---- lists: shared rows -> global, chunk-transposed: chunk ch of particle i is the uint4 (t * 13 + ch) * 32 + lane
        __syncwarp();
        {
            const int nch = live ? (cnt + 7) >> 3 : 0;
            const int maxch = __reduce_max_sync(BBX_TP_FULL, nch);
            uint4 *dst = reinterpret_cast<uint4 *>(nbr) + ((size_t)t * BBX_NBR_CHUNKS) * 32 + lane;
            const uint4 *src = reinterpret_cast<const uint4 *>(my_row);
            for(int ch = 0; ch < maxch; ch++) if(ch < nch) dst[(size_t)ch * 32] = src[ch];
        }
        __syncwarp();
    }
}
