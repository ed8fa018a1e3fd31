// bbx list build, v12 (phase B): per-particle neighbour lists + density, ONE THREAD PER PARTICLE, one warp per TILE of 32
// consecutive slots, the candidates streamed through shared memory run by run.
//
// Why this shape (profiles/r02_notes.md): the warp-per-cell kernel (bbx_lists.cuh, v7) spends 16 instructions per
// (32 candidates, 1 particle) in its pair loop -- and as much again staging ~313 candidates per cell, of which 2/3 are
// re-staged by the next cell of the x-row, and copying 4 list rows out per group.  Here
//   * slots are sorted by cell and x is the fastest cell index, so the 32 particles of a tile sit in ~4 consecutive cells of
//     one x-row and, for each of the 9 (dy, dz) offsets ("runs"), everything they can see is ONE contiguous slot range of
//     ~60 particles: one bulk copy per run (cp.async.bulk + mbarrier, UBLKCP in SASS) brings the raw positions in, the copy
//     of run r + 1 in flight while run r is walked (two buffers) -- ~18 candidates staged per particle instead of ~313 per
//     10, no per-candidate copy instruction, 4 KB of stage per warp;
//   * every lane walks ITS OWN 3-cell window of the run in lock step with the others (lanes of one cell read the same
//     address: a broadcast), so the list order is the walk order and the 100 cap and the FP64 band check are thread-local;
//   * accepted entries collect in a small per-lane buffer; after every run its complete 8-entry chunks go straight to the
//     lane's place in the chunk-transposed block of `nbr` (the tile's 32 rows ARE one such block);
//   * the loop over the runs is rolled and everything per-run lives in shared-memory tables: the v11 kernel, unrolled 9
//     times with ~75 KB of stage per CTA, was bound by its instruction-cache misses and its 12 warps per SM.
// A tile that spans two x-rows runs one pass per row; a group of lanes whose run does not fit the stage is halved until it
// does; a single particle whose own run exceeds it, and a particle that runs into the 100 cap, is redone by
// bbx_list_walk_exact straight from global memory in the reference's traversal order.
//
// Arithmetic: d^2 straight from the raw FP32 positions in one fixed operation order, so every value (acceptance, density
// weight) depends on the pair alone and not on the tile it was computed in: a slab engine, whose slots are numbered
// differently, reproduces the single-domain engine bit for bit.  d^2 < thr_lo: certainly inside; thr_lo <= d^2 <= thr_hi (the
// guard band around h^2 - 1e-8): appended provisionally, and a lane that took such a candidate redoes the run with the
// reference's FP64 predicate deciding (IsWithinStd, kernel.cpp:229-234).
// Reference: Grid::DistributeParticleBucket grid.h:422-447, Bucket::Insert particle.h:44-50 (cap 100),
// ComputeDensityFor + ComputePressureValue sph_equations3.cpp:7-58.
#pragma once
#include "bbx_device.cuh"

#define BBX_TP_WARPS 4            // warps (tiles in flight) per CTA
#ifndef BBX_TP_RCAP
#define BBX_TP_RCAP 128           // staged candidates per run buffer (16 B each, two buffers per warp); <= 255
#endif
#ifndef BBX_TP_LB
#define BBX_TP_LB 56              // u16 entries of a lane's list buffer (multiple of 8; 112 B rows: conflict-free LDS.128)
#endif
#ifndef BBX_TP_UNROLL
#define BBX_TP_UNROLL 4
#endif
#define BBX_TP_OFF_TAB (2 * BBX_TP_RCAP * 16)                 // u16 [9][32]: (offset in the run buffer) | (window length) << 8
#define BBX_TP_OFF_UTAB (BBX_TP_OFF_TAB + 9 * 32 * 2)         // int [9][2]: first slot and length of the run of the group
#define BBX_TP_OFF_LIST (BBX_TP_OFF_UTAB + 80)
#define BBX_TP_OFF_MBAR (BBX_TP_OFF_LIST + 32 * BBX_TP_LB * 2)
#define BBX_TP_WARP_BYTES (BBX_TP_OFF_MBAR + 16)
#define BBX_TP_SMEM (BBX_TP_WARPS * BBX_TP_WARP_BYTES)
#ifndef BBX_TP_MINB
#define BBX_TP_MINB 6
#endif
#define BBX_TP_FULL 0xffffffffu

// ---- mbarrier / bulk-copy plumbing (PTX; SASS: SYNCS.*, UBLKCP)
__device__ __forceinline__ unsigned bbx_smem_u32(const void *p){ return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bbx_mbar_init(unsigned long long *b, unsigned count){
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bbx_smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void bbx_mbar_arrive_expect_tx(unsigned long long *b, unsigned bytes){
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bbx_smem_u32(b)), "r"(bytes) : "memory");
}
// bounded: a copy that never lands (it cannot, short of a driver fault) becomes the sticky device error BBX_ERR_COMM after
// ~2^26 polls instead of a hang; returns false in that case
__device__ __forceinline__ bool bbx_mbar_wait(unsigned long long *b, unsigned parity){
    unsigned ok = 0;
    for(int spin = 0; spin < (1 << 26); spin++){
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bbx_smem_u32(b)), "r"(parity) : "memory");
        if(ok) return true;
    }
    return false;
}
// global -> shared, `bytes` (multiple of 16, both addresses 16-byte aligned), completion counted on the mbarrier
__device__ __forceinline__ void bbx_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *b){
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(bbx_smem_u32(dst)), "l"(src), "r"(bytes), "r"(bbx_smem_u32(b)) : "memory");
}

__device__ __forceinline__ uint4 bbx_lds128(unsigned addr){
    uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)); return v;
}
__device__ __forceinline__ void bbx_sts128(unsigned addr, uint4 v){
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// One particle straight from global memory, in the reference's traversal order (y outer, x middle, z inner; chain order
// inside a cell): keeps the first 100 accepted exactly like Bucket::Insert, the density is the sum over exactly those;
// entries go straight to the particle's place in the chunk-transposed `nbr`.
// Used for a particle that ran into the cap and for one whose runs do not fit the stage.
__device__ __noinline__ void bbx_list_walk_exact(const StepParams &P, const DevGrid &g, DevState *st,
        const float4 *__restrict__ pos, const int *__restrict__ cell_start, unsigned short *__restrict__ nbr,
        int i, int c, float4 pi, int *cnt_out, float *sum_out)
{
    int cnt = 0, total = 0; float sum = 0.f;
    int cz = c / g.plane; int rem = c - cz * g.plane; int cy = rem / g.n[0]; int cx = rem - cy * g.n[0];
    int xlo = max(cx - 1, 0);
    unsigned short *row = nbr + ((size_t)(i >> 5) * BBX_NBR_CHUNKS * 32 + (i & 31)) * 8;   // entry k: row[(k >> 3) * 256 + (k & 7)]
    for(int dy = -1; dy <= 1; dy++) for(int dx_ = -1; dx_ <= 1; dx_++) for(int dz = -1; dz <= 1; dz++){
        int x = cx + dx_, y = cy + dy, z = cz + dz;
        if(x < 0 || x >= g.n[0] || y < 0 || y >= g.n[1] || z < 0 || z >= g.n[2]) continue;
        int nb = x + y * g.n[0] + z * g.plane;
        int r = (dy + 1) * 3 + (dz + 1);
        int rb = cell_start[xlo + y * g.n[0] + z * g.plane];
        int s = cell_start[nb], e = cell_start[nb + 1];
        for(int j = s; j < e; j++){
            if(j - rb >= BBX_MAX_RUN_LEN){ st->error = BBX_ERR_CAPACITY; break; }
            float4 pj = pos[j];
            float ddx = pi.x - pj.x, ddy = pi.y - pj.y, ddz = pi.z - pj.z;
            float d2 = fmaf(ddx, ddx, fmaf(ddy, ddy, ddz * ddz));
            if(bbx_accept(P, pi, pj, d2)){
                total++;
                if(cnt < BBX_MAX_NEIGHBORS){
                    const float yy = P.h2 - d2;
                    sum = fmaf(yy * yy, yy, sum);
                    row[(cnt >> 3) * 256 + (cnt & 7)] = (unsigned short)(((unsigned)r << BBX_RUN_SHIFT) | (unsigned)(j - rb));
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
    float4 *buf = reinterpret_cast<float4 *>(wbase);                                        // [2][BBX_TP_RCAP] raw positions of a run
    unsigned short *tab = reinterpret_cast<unsigned short *>(wbase + BBX_TP_OFF_TAB);
    int *utab = reinterpret_cast<int *>(wbase + BBX_TP_OFF_UTAB);
    const unsigned lbase = bbx_smem_u32(wbase + BBX_TP_OFF_LIST + lane * (BBX_TP_LB * 2)); // this lane's list buffer
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(wbase + BBX_TP_OFF_MBAR);   // one per run buffer
    if(lane == 0){ bbx_mbar_init(mbar, 1); bbx_mbar_init(mbar + 1, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    unsigned par = 0;                  // bit b: phase parity the next wait on mbar[b] expects
    const int n = bbx_count(P);
    H = bbx_halo_resolve(H, st);
    const PartMap pm = bbx_part_map(P, 32, n);          // (slab engines: boundary tiles first, see bbx_part_map)
    const int nvt = bbx_part_blocks(P, pm), nwarps = gridDim.x * BBX_TP_WARPS;
#pragma unroll 1
    for(int vt = blockIdx.x * BBX_TP_WARPS + warp; vt < nvt; vt += nwarps){
        const int t = bbx_part_block(P, pm, vt);
        const int i = t * 32 + lane;
        const bool live = i < n;
        int c = 0; float4 pi = make_float4(0.f, 0.f, 0.f, 0.f);
        if(live){ c = cell[i]; pi = pos[i]; }
        const int rowid = c / g.n[0], cx = c - rowid * g.n[0];
        const int cz = rowid / g.n[1], cy = rowid - cz * g.n[1];
        uint4 *dst = reinterpret_cast<uint4 *>(nbr) + ((size_t)t * BBX_NBR_CHUNKS) * 32 + lane;   // chunk ch of this lane: dst[ch * 32]
        unsigned wa = lbase;            // shared-memory address of the next list entry of this lane
        int fl = 0;                     // chunks of this lane already in `nbr`
        float acc = 0.f;
        bool slow = false;
        float4 cj[BBX_TP_UNROLL];       // (a lane past the end of its window keeps stale, finite values: never accepted)
#pragma unroll
        for(int u = 0; u < BBX_TP_UNROLL; u++) cj[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned todo = __ballot_sync(BBX_TP_FULL, live);
#pragma unroll 1
        while(todo){
            // ---- the group of this pass: the remaining lanes of the first remaining lane's x-row, halved until every run fits;
            // the run tables of the group go to shared memory (the loop over the runs below is rolled)
            const int f = __ffs(todo) - 1;
            const int row_f = __shfl_sync(BBX_TP_FULL, rowid, f);
            const int cyf = __shfl_sync(BBX_TP_FULL, cy, f), czf = __shfl_sync(BBX_TP_FULL, cz, f);
            unsigned grp = __ballot_sync(BBX_TP_FULL, live && rowid == row_f) & todo;
            bool fits, member;
            int T, l;
#pragma unroll 1
            for(;;){
                l = 31 - __clz(grp);
                member = (grp >> lane) & 1u;
                fits = true; T = 0;
                const int xlo = max(cx - 1, 0), xhi = min(cx + 1, g.n[0] - 1);
                __syncwarp();
#pragma unroll
                for(int r = 0; r < 9; r++){
                    const int y = cyf + r / 3 - 1, z = czf + r % 3 - 1;
                    int lo = 0, hi = 0;
                    if(y >= 0 && y < g.n[1] && z >= 0 && z < g.n[2]){
                        const int rb = y * g.n[0] + z * g.plane;
                        lo = cell_start[rb + xlo]; hi = cell_start[rb + xhi + 1];
                    }
                    const int a = __shfl_sync(BBX_TP_FULL, lo, f), b = __shfl_sync(BBX_TP_FULL, hi, l);
                    if(b - a > BBX_TP_RCAP) fits = false;
                    tab[r * 32 + lane] = member ? (unsigned short)((unsigned)(lo - a) | ((unsigned)(hi - lo) << 8)) : (unsigned short)0;
                    if(lane == r){ utab[2 * r] = a; utab[2 * r + 1] = b - a; }
                    T += hi - lo;
                }
                if(fits || f == l) break;
                grp = ((1u << ((__popc(grp) + 1) >> 1)) - 1u) << f;   // (the lanes of a group are contiguous)
            }
            todo &= ~grp;
            {
                const int tmax = __reduce_max_sync(BBX_TP_FULL, member ? T : 0);
                if(lane == 0 && tmax > st->max_candidates) atomicMax(&st->max_candidates, tmax);
            }
            if(!fits){ if(lane == f) slow = true; continue; }          // one particle whose own run is longer than the stage
            // ---- the runs, streamed: the copy of run r + 1 is in flight while run r is walked
#define BBX_TP_ISSUE(q)                                                                                                     \
            {                                                                                                               \
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   /* earlier generic-proxy reads of the buffer */ \
                __syncwarp();                                                                                               \
                if(lane == 0){                                                                                              \
                    const int a_ = utab[2 * (q)], ul_ = utab[2 * (q) + 1];                                                  \
                    bbx_mbar_arrive_expect_tx(mbar + ((q) & 1), 16u * (unsigned)ul_);                                       \
                    if(ul_ > 0) bbx_bulk_g2s(buf + ((q) & 1) * BBX_TP_RCAP, pos + a_, 16u * (unsigned)ul_, mbar + ((q) & 1)); \
                }                                                                                                           \
            }
            BBX_TP_ISSUE(0)
#pragma unroll 1
            for(int r = 0; r < 9; r++){
                if(r < 8) BBX_TP_ISSUE(r + 1)
                if(!bbx_mbar_wait(mbar + (r & 1), (par >> (r & 1)) & 1u)) st->error = BBX_ERR_COMM;
                par ^= 1u << (r & 1);
                const unsigned e = tab[r * 32 + lane];
                const int mylen = (int)(e >> 8);                         // (0 outside the group)
                const float4 *cp = buf + (r & 1) * BBX_TP_RCAP + (e & 255u);
                const int L = __reduce_max_sync(BBX_TP_FULL, mylen);
                const int Lall = __reduce_min_sync(BBX_TP_FULL, member ? mylen : 0x7fffffff) & ~(BBX_TP_UNROLL - 1);
                const unsigned lim = lbase + 2u * (BBX_TP_LB - BBX_TP_UNROLL);
                const unsigned wa0 = wa; const float acc0 = acc;
                const unsigned ent = (unsigned)r << BBX_RUN_SHIFT;
                float dmx = 0.f;
                int k0 = 0;
                const float thr_m = member ? P.thr_hi : -1.f;
                // A candidate up to thr_hi is appended provisionally; a lane that took one from inside the guard band redoes
                // the run with the FP64 predicate.  Blocks inside every member's window run without the per-step window test.
// one pair test + predicated append, in PTX so that it stays four predicated instructions (the compiler otherwise turns
// the append into a branch that ~every warp takes): THR is thr_hi for a lane inside its window and -1 outside
#define BBX_TP_STEP(THR)                                                                                                     \
                    {                                                                                                        \
                        const float dx = pi.x - cj[u].x, dy = pi.y - cj[u].y, dz = pi.z - cj[u].z;                           \
                        const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));                                                \
                        const float y = P.h2 - d2, y2 = y * y;                                                               \
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %3, %4;\n\t@p st.shared.u16 [%0], %5;\n\t@p add.u32 %0, %0, 2;\n\t" \
                                     "@p fma.rn.f32 %1, %6, %7, %1;\n\t@p max.f32 %2, %2, %3;\n\t}"                      \
                                     : "+r"(wa), "+f"(acc), "+f"(dmx)                                                        \
                                     : "f"(d2), "f"(THR), "h"((unsigned short)(ent + (unsigned)(k0 + u))), "f"(y2), "f"(y) : "memory"); \
                    }
#pragma unroll 1
                for(; k0 < Lall; k0 += BBX_TP_UNROLL){
                    // the buffer holds BBX_TP_LB entries: a lane that would overrun it is redone by the slow path
                    if(wa > lim){ slow = true; wa = lbase; }
#pragma unroll
                    for(int u = 0; u < BBX_TP_UNROLL; u++) cj[u] = cp[k0 + u];
#pragma unroll
                    for(int u = 0; u < BBX_TP_UNROLL; u++) BBX_TP_STEP(thr_m)
                }
#pragma unroll 1
                for(; k0 < L; k0 += BBX_TP_UNROLL){
                    if(wa > lim){ slow = true; wa = lbase; }
#pragma unroll
                    for(int u = 0; u < BBX_TP_UNROLL; u++) if(k0 + u < mylen) cj[u] = cp[k0 + u];
#pragma unroll
                    for(int u = 0; u < BBX_TP_UNROLL; u++) BBX_TP_STEP(k0 + u < mylen ? P.thr_hi : -1.f)
                }
#undef BBX_TP_STEP
                if(dmx >= P.thr_lo){
                    wa = wa0; acc = acc0;
#pragma unroll 1
                    for(int k = 0; k < mylen; k++){
                        if(wa > lim){ slow = true; wa = lbase; }
                        const float4 c4 = cp[k];
                        const float dx = pi.x - c4.x, dy = pi.y - c4.y, dz = pi.z - c4.z;
                        const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        bool in = d2 < P.thr_lo;
                        if(!in && d2 <= P.thr_hi){ in = bbx_within_std_exact(pi, c4, P.h2_d); atomicAdd(&st->exact_passes, 1); }
                        if(in){
                            const float y = P.h2 - d2;
                            asm volatile("st.shared.u16 [%0], %1;" :: "r"(wa), "h"((unsigned short)(ent + (unsigned)k)) : "memory");
                            wa += 2u;
                            acc = fmaf(y * y, y, acc);
                        }
                    }
                }
                // ---- complete chunks of the buffer -> the lane's row of `nbr`; the partial one moves to the front
                {
                    int nb = (int)(wa - lbase) >> 1;
                    // (wa > lim: the block-start test of the walk, applied once more at the end of the run -- whether the warp
                    // ran further blocks after this lane's window depends on the OTHER lanes of the tile, the outcome must not)
                    if(8 * fl + nb > BBX_MAX_NEIGHBORS || wa > lim) slow = true;     // (the slow path rewrites the whole row)
                    if(slow){ wa = lbase; nb = 0; }
                    const int nfull = nb >> 3;
                    const int mx = __reduce_max_sync(BBX_TP_FULL, nfull);
                    for(int ch = 0; ch < mx; ch++) if(ch < nfull) dst[(size_t)(fl + ch) * 32] = bbx_lds128(lbase + 16u * (unsigned)ch);
                    if(nfull){
                        const uint4 rest = bbx_lds128(lbase + 16u * (unsigned)nfull);
                        bbx_sts128(lbase, rest);
                        wa -= 16u * (unsigned)nfull;
                        fl += nfull;
                    }
                }
            }
#undef BBX_TP_ISSUE
        }
        // ---- per particle: the last partial chunk, count, density, the records of the force sweeps
        int cnt = 8 * fl + ((int)(wa - lbase) >> 1);
        float sum = acc;
        if(!slow && wa > lbase) dst[(size_t)fl * 32] = bbx_lds128(lbase);
        if(slow) bbx_list_walk_exact(P, g, st, pos, cell_start, nbr, i, c, pi, &cnt, &sum);
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
        __syncwarp();
    }
}
