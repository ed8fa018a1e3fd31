// bbx list build (phase B, engine v8): per-particle neighbour lists + density, ONE WARP PER OCCUPIED CELL,
// "transposed": the 32 lanes hold 32 consecutive CANDIDATES of the cell's 27-cell neighbourhood, the own
// particles of the cell are held in REGISTERS, 4 at a time.
//
// History (ncu, profiles/): one thread per particle diverges on the candidate walk (v2); v4/v5 made the lanes
// candidates and looped over 16 own particles broadcast from shared memory -- every pair test then started
// with an LDS.128 whose latency the 4-5 resident warps per scheduler could not hide.  v6/v7 staged the cell's
// candidates once per cell in shared memory (cell frame, culled against the cell's box) and kept the own
// particles in registers; ncu then showed 32 % of the kernel's instructions in the STAGING: every warp fetched
// its ~313 candidates from global memory one by one (a cp.async per candidate behind a run search), although
// x-neighbouring cells share two thirds of them.  v8 (this file):
//   * a CTA (4 warps) works on a SUB-CHUNK of up to 4 consecutive occupied cells of one x-row; the union of their
//     neighbourhoods is 9 contiguous slot ranges (one per (dy, dz): x is the fastest cell index), which ONE elected
//     warp moves into a shared POOL with 9 bulk copies (cp.async.bulk.shared.global + mbarrier complete_tx: SASS
//     UBLKCP) -- no per-candidate instruction, each candidate crosses the L2 once per sub-chunk instead of 3 times;
//   * no __syncthreads in the loop: the pool of the next sub-chunk is requested by whichever warp is the LAST to
//     finish reading the current one (a shared counter), so the copy runs under the other warps' pair loops and a
//     warp only ever waits on the mbarrier of a copy issued long before;
//   * each warp then compacts ITS cell's window of the pool into a private array in the cell frame
//     u = (x - cell centre) / h (the subtraction is exact in FP32), dropping candidates farther than h from the
//     cell's box (-24 %), list entry (run, offset) beside them;
//   * a group of 4 own particles lives in registers as (2 u_i, 1 - |u_i|^2); the round loop reads one
//     candidate per lane (conflict-free LDS.128, fetched one round ahead) and runs the pair tests on registers;
//   * the pair test is x = 1 - d^2 / h^2 = (1 - |u_i|^2) - |u_j|^2 + 2 u_j . u_i : 3 FFMA + 1 FADD; accepted when
//     x > xacc, and the same x is the density weight (W_std ~ x^3);
//   * the accept decisions of (round, particle) are ONE ballot; an accepted lane's list position is the
//     particle's running count + popc(ballot & lanes below) -- no atomics, deterministic flat order.
// ~16 instructions per (round, particle).
//
// Candidates are visited in the flat order run-major / slot order and lists keep the format of v2:
// entry = (run << 12) | offset-in-run, chunk-transposed per 32 particles (the sweeps' coalesced 512 B loads).
// A sub-chunk whose pool would not fit (compressed cells) is shortened; a single cell whose neighbourhood does
// not fit the pool or whose culled candidates do not fit the warp's array is staged from global memory in
// several pieces per group (rare path, bbx_list_stage_global).
//
// Exactness: in the cell frame |u| <= ~2.6, so x carries an absolute error of a few 1e-7.  A candidate is
// accepted when x > xacc (d^2 certainly below h^2 - 1e-8 + band).  If any accepted candidate of a group lay
// inside the guard band (x < xband), the group is redone with the FP64 predicate of the reference
// (IsWithinStd, kernel.cpp:229-234) deciding the band members -- ~1 % of the groups.
// Reference: Grid::DistributeParticleBucket grid.h:422-447, Bucket::Insert particle.h:44-50 (cap 100),
// ComputeDensityFor + ComputePressureValue sph_equations3.cpp:7-58.
#pragma once
#include "bbx_device.cuh"

#define BBX_LW 4                  // warps per CTA = cells per sub-chunk
#define BBX_LT (BBX_LW * 32)
#ifndef BBX_G
#define BBX_G 4                   // own particles per group (registers)
#endif
#define BBX_G_SHIFT (BBX_G == 8 ? 2 : 3) // lane >> shift = particle whose density total the lane holds after the butterfly
#ifndef BBX_CMAX
#define BBX_CMAX 384              // culled candidates a warp stages at a time (multiple of 32; a dense cell keeps ~240)
#endif
#ifndef BBX_POOL
#define BBX_POOL 768              // float4 slots of the CTA's candidate pool (4 dense cells need 9 x 6 x 11.7 = 632)
#endif
#define BBX_ROW 104               // u16 entries per list row in shared memory (13 chunks of 8)
#define BBX_FULL 0xffffffffu
#ifndef BBX_LIST_MINB
#define BBX_LIST_MINB 5           // resident CTAs per SM the list kernel is compiled for
#endif

// ---- mbarrier / bulk-copy plumbing (PTX; SASS: SYNCS.*, UBLKCP)
__device__ __forceinline__ unsigned bbx_smem_u32(const void *p){ return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bbx_mbar_init(unsigned long long *b, unsigned count){
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bbx_smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void bbx_mbar_arrive(unsigned long long *b){
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bbx_smem_u32(b)) : "memory");
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

// Slow path of one particle whose list would exceed 100 entries: re-walk the 27 cells in the reference's
// order (y outer, x middle, z inner; chain order inside a cell) and keep the first 100 exactly like
// Bucket::Insert; the density is the sum over exactly those.
__device__ __noinline__ void bbx_list_overflow_row(const StepParams &P, const DevGrid &g, DevState *st,
        const float4 *__restrict__ pos, const int *__restrict__ cell_start, unsigned short *row,
        int c, float4 pi, int *cnt_out, float *sum_out)
{
    atomicAdd(&st->overflow, 1);
    int cnt = 0; float sum = 0.f;
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
                row[cnt] = (unsigned short)(((unsigned)r << BBX_RUN_SHIFT) | (unsigned)(j - rb));
                cnt++;
            }
        }
    }
    *cnt_out = cnt; *sum_out = sum;
}

struct ListWarp {
    float4 *cand;            // [BBX_CMAX + 32] staged candidates (u_j.x, u_j.y, u_j.z, bits of the list entry)
    float4 *spi;             // [BBX_G] own particles of the group, raw positions (FP64 re-check, cap-100 slow path)
    unsigned short *rows;    // [BBX_G][BBX_ROW] lists being built
    int *tab;                // [0..9] exclusive prefix of the 9 window lengths (tab[9] = T), [10..18] first GLOBAL slot of each window,
                             // [19..27] POOL index of the first slot of each window
};
// the cell frame of the pair test and the half extents of the cell in it (culling)
struct ListFrame { float cx, cy, cz, hx, hy, hz; };

// cell frame + cull of one raw candidate: kept when closer than h to the cell's box (it could be a neighbour of
// SOME particle of the cell)
__device__ __forceinline__ bool bbx_list_frame(const StepParams &P, const ListFrame &F, float4 raw, float &ux, float &uy, float &uz){
    ux = (raw.x - F.cx) * P.inv_h; uy = (raw.y - F.cy) * P.inv_h; uz = (raw.z - F.cz) * P.inv_h;
    const float gx = fmaxf(fabsf(ux) - F.hx, 0.f), gy = fmaxf(fabsf(uy) - F.hy, 0.f), gz = fmaxf(fabsf(uz) - F.hz, 0.f);
    return fmaf(gx, gx, fmaf(gy, gy, gz * gz)) < 1.001f;
}
__device__ __forceinline__ void bbx_list_pad(const ListWarp &W, int n, int lane){
    __syncwarp();
    if(n + lane < ((n + 31) & ~31)) W.cand[n + lane] = make_float4(1.0e15f, 0.f, 0.f, 0.f); // never accepted
    __syncwarp();
}

// Hot path: the warp's window of the CTA's pool -> its private candidate array, flat order (run-major, slot order),
// cell frame, culled (ballot compaction keeps the order).  Returns the number kept, or -1 when they do not fit
// BBX_CMAX (the cell then takes the piecewise path).
__device__ __forceinline__ int bbx_list_compact_pool(const StepParams &P, DevState *st, const ListWarp &W, const float4 *pool,
        int T, int lane, const ListFrame &F)
{
    const unsigned lt = lanemask_lt();
    int n = 0, r = 0;
#pragma unroll 1
    for(int k0 = 0; k0 < T; k0 += 32){
        const int f = min(k0 + lane, T - 1);
        // windows only move forward (tab[9] = T > f): 32 candidates further on a lane is at most two non-empty windows on
        // in a dense neighbourhood -- two predicated steps, then the general loop for what is left (sparse cells)
        r += (r < 8 && f >= W.tab[r + 1]) ? 1 : 0;
        r += (r < 8 && f >= W.tab[r + 1]) ? 1 : 0;
        while(f >= W.tab[r + 1]) r++;
        const int off = f - W.tab[r];
        const float4 raw = pool[W.tab[19 + r] + off];
        float ux, uy, uz;
        const bool keep = bbx_list_frame(P, F, raw, ux, uy, uz) && (k0 + lane < T);
        if(off >= BBX_MAX_RUN_LEN) st->error = BBX_ERR_CAPACITY;
        const unsigned entry = ((unsigned)r << BBX_RUN_SHIFT) | (unsigned)min(off, BBX_MAX_RUN_LEN - 1);
        const unsigned msk = __ballot_sync(BBX_FULL, keep);
        const int at = n + __popc(msk & lt);
        if(keep && at < BBX_CMAX) W.cand[at] = make_float4(ux, uy, uz, __uint_as_float(entry));
        n += __popc(msk);
    }
    if(n > BBX_CMAX) return -1;
    bbx_list_pad(W, n, lane);
    return (n + 31) & ~31; // whole rounds (the padding is never accepted)
}

// Rare path (neighbourhood too large for the pool / the warp's array): candidates [f0, ...) of the flat order straight
// from global memory, up to BBX_CMAX kept per call; f0 advances (f0 = T when all are consumed).
__device__ __noinline__ int bbx_list_stage_global(const StepParams &P, DevState *st, const ListWarp &W, int T, int lane, int *f0_io,
        const ListFrame &F, const float4 *__restrict__ pos)
{
    __syncwarp();
    const unsigned lt = lanemask_lt();
    int n = 0, r = 0, k0 = *f0_io;
    while(k0 < T && n + 32 <= BBX_CMAX){
        const int f = min(k0 + lane, T - 1);
        while(f >= W.tab[r + 1]) r++;
        const int off = f - W.tab[r];
        const float4 raw = pos[W.tab[10 + r] + off];
        float ux, uy, uz;
        const bool keep = bbx_list_frame(P, F, raw, ux, uy, uz) && (k0 + lane < T);
        if(off >= BBX_MAX_RUN_LEN) st->error = BBX_ERR_CAPACITY;
        const unsigned entry = ((unsigned)r << BBX_RUN_SHIFT) | (unsigned)min(off, BBX_MAX_RUN_LEN - 1);
        const unsigned msk = __ballot_sync(BBX_FULL, keep);
        if(keep) W.cand[n + __popc(msk & lt)] = make_float4(ux, uy, uz, __uint_as_float(entry));
        n += __popc(msk);
        k0 += 32;
    }
    *f0_io = min(k0, T);
    bbx_list_pad(W, n, lane);
    return n;
}

// All staged candidates against NG (2 or 4) own particles of the group in registers.  Accumulates cnt / acc over
// pieces; xmin = smallest x among this lane's accepted pairs.  The candidate of the NEXT round is in flight while the
// current one is tested (two register sets, the loop body is written out twice: no copies, no address rebuild).
__device__ __forceinline__ float4 bbx_lds128(unsigned addr){
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
template<bool EXACT, int NG>
__device__ __forceinline__ void bbx_list_rounds(const StepParams &P, const ListWarp &W, int nc, int lane,
        const float4 (&q)[BBX_G], const float4 *__restrict__ pos, int (&cnt)[BBX_G], float (&acc)[BBX_G], float &xmin)
{
    if(nc <= 0) return;
    const unsigned lt = lanemask_lt();
    // shared-memory byte addresses pinned in registers (the compiler otherwise rebuilds them from special registers
    // in front of every access)
    unsigned rows_addr, ca;
    asm volatile("mov.u32 %0, %1;" : "=r"(rows_addr) : "r"(bbx_smem_u32(W.rows)));
    asm volatile("mov.u32 %0, %1;" : "=r"(ca) : "r"(bbx_smem_u32(W.cand) + 16u * (unsigned)lane));
    const unsigned ca_end = ca + 16u * (unsigned)nc; // nc is a multiple of 32 (padded)
    float xacc;
    asm volatile("mov.f32 %0, %1;" : "=f"(xacc) : "f"(P.xacc));
    auto round = [&](const float4 cj){
        const float a = fmaf(cj.x, cj.x, fmaf(cj.y, cj.y, cj.z * cj.z));
        const unsigned short entry = (unsigned short)__float_as_uint(cj.w);
#pragma unroll
        for(int ii = 0; ii < NG; ii++){
            const float x = fmaf(cj.x, q[ii].x, fmaf(cj.y, q[ii].y, fmaf(cj.z, q[ii].z, q[ii].w))) - a;
            bool in = x > xacc;
            if(EXACT){
                if(in && x < P.xband){
                    const unsigned e = entry;
                    in = bbx_within_std_exact(W.spi[ii], pos[W.tab[10 + (e >> BBX_RUN_SHIFT)] + (int)(e & BBX_RUN_MASK)], P.h2_d);
                }
            }else{
                if(in) xmin = fminf(xmin, x);
            }
            const unsigned msk = __ballot_sync(BBX_FULL, in);
            // list position = entries so far + accepted lanes below this one (flat order); rows have
            // 104 slots, the count keeps running so that the cap-100 slow path can be detected
            const int k = min(cnt[ii] + __popc(msk & lt), BBX_ROW - 1);
            if(in) asm volatile("st.shared.u16 [%0], %1;" :: "r"(rows_addr + (unsigned)(ii * BBX_ROW * 2) + 2u * (unsigned)k), "h"(entry) : "memory");
            cnt[ii] += __popc(msk);
            const float x2 = x * x;
            acc[ii] = in ? fmaf(x2, x, acc[ii]) : acc[ii];
        }
    };
    float4 c0 = bbx_lds128(ca), c1 = c0;
#pragma unroll 1
    for(;;){
        if(ca + 512u < ca_end) c1 = bbx_lds128(ca + 512u);
        round(c0);
        if(ca + 512u >= ca_end) break;
        if(ca + 1024u < ca_end) c0 = bbx_lds128(ca + 1024u);
        round(c1);
        ca += 1024u;
        if(ca >= ca_end) break;
    }
}

// One group: the whole neighbourhood against the group's particles (nc >= 0: it is staged already, nc candidates;
// else piece by piece from global memory).  Returns (warp-uniform) whether a provisionally accepted candidate
// lay inside the guard band.
template<bool EXACT>
__device__ __forceinline__ bool bbx_list_group(const StepParams &P, DevState *st, const ListWarp &W, int T, int nc, int mg, int lane,
        const ListFrame &F, const float4 (&q)[BBX_G], const float4 *__restrict__ pos, int (&cnt)[BBX_G], float (&acc)[BBX_G])
{
#pragma unroll
    for(int ii = 0; ii < BBX_G; ii++){ acc[ii] = 0.f; cnt[ii] = 0; }
    float xmin = 1.0e30f;
    if(nc >= 0){
        // a tail group of 1 or 2 particles tests 2 slots instead of 4
        if(mg > 2) bbx_list_rounds<EXACT, BBX_G>(P, W, nc, lane, q, pos, cnt, acc, xmin);
        else bbx_list_rounds<EXACT, 2>(P, W, nc, lane, q, pos, cnt, acc, xmin);
    }else{
        int f0 = 0;
#pragma unroll 1
        while(f0 < T){
            const int n = bbx_list_stage_global(P, st, W, T, lane, &f0, F, pos);
            bbx_list_rounds<EXACT, BBX_G>(P, W, (n + 31) & ~31, lane, q, pos, cnt, acc, xmin);
        }
    }
    return __any_sync(BBX_FULL, xmin < P.xband);
}

// What the CTA shares: the pool, the two descriptor slots of the sub-chunk pipeline, the warps' private arrays.
struct __align__(16) ListShared {
    float4 pool[BBX_POOL];
    float4 cand[BBX_LW][BBX_CMAX + 32];
    float4 spi[BBX_LW][BBX_G];
    unsigned short rows[BBX_LW][BBX_G * BBX_ROW];
    unsigned long long full[2];   // mbarrier per descriptor slot: descriptors written and pool copy complete
    int tab[BBX_LW][28];
    float ovs[BBX_LW][BBX_G];
    int sc_cell[2][BBX_LW];       // cell of warp w in the sub-chunk (-1: none)
    int sc_base[2][9];            // first global slot of pool run r
    int sc_off[2][10];            // pool index of run r (exclusive prefix), [9] = total
    int sc_flag[2];               // 0: pool staged, 1: no more work, 2: pool not staged (the cell takes the piecewise path)
    int done;                     // warps that have finished reading the pool of the current sub-chunk
    int next_q, next_a;           // producer state: next chunk of 4 occupied cells / first unprocessed cell of it
};

// Producer (one whole warp): describe the next sub-chunk in descriptor slot `slot` and start its pool copy.
// A sub-chunk = the longest prefix of the remaining cells of chunk q (4 consecutive entries of the occupied-cell
// list) that lies in ONE x-row and whose 9 union ranges fit the pool.
__device__ __forceinline__ void bbx_list_produce(ListShared &S, int slot, const DevGrid &g, int n_occ, const int *__restrict__ occ_cells,
        const int *__restrict__ cell_start, const float4 *__restrict__ pos, int lane)
{
    const int q = S.next_q, a = S.next_a;
    if(q * BBX_LW + a >= n_occ){
        if(lane == 0){ S.sc_flag[slot] = 1; bbx_mbar_arrive(&S.full[slot]); }
        return;
    }
    const int clen = min(BBX_LW, n_occ - q * BBX_LW);
    const int ct = lane < clen ? occ_cells[q * BBX_LW + lane] : -1;
    const int ca = __shfl_sync(BBX_FULL, ct, a);
    const int rowa = ca / g.n[0];
    // cells a, a+1, ... of the chunk that share the row of cell a
    const unsigned same = __ballot_sync(BBX_FULL, lane >= a && lane < clen && ct / g.n[0] == rowa) >> a;
    int b = a + (__ffs(~same) - 1);
    const int cz = ca / g.plane, cy = rowa - cz * g.n[1], xa = ca - rowa * g.n[0];
    int base = 0, len = 0, inc = 0, total = 0;
    for(;;){
        const int xb = __shfl_sync(BBX_FULL, ct, b - 1) - rowa * g.n[0];
        base = 0; len = 0;
        if(lane < 9){
            const int y = cy + lane / 3 - 1, z = cz + lane % 3 - 1;
            if(y >= 0 && y < g.n[1] && z >= 0 && z < g.n[2]){
                const int row = y * g.n[0] + z * g.plane;
                base = cell_start[row + max(xa - 1, 0)];
                len = cell_start[row + min(xb + 1, g.n[0] - 1) + 1] - base;
            }
        }
        inc = len;
#pragma unroll
        for(int o = 1; o < 16; o <<= 1){ const int y = __shfl_up_sync(BBX_FULL, inc, o); if(lane >= o) inc += y; }
        total = __shfl_sync(BBX_FULL, inc, 15);
        if(total <= BBX_POOL || b == a + 1) break;
        b--;
    }
    const bool staged = total <= BBX_POOL;
    const int cw = __shfl_sync(BBX_FULL, ct, min(a + lane, BBX_LW - 1)); // warp w works on cell a + w of the chunk
    if(lane < BBX_LW) S.sc_cell[slot][lane] = (a + lane < b) ? cw : -1;
    if(lane < 10) S.sc_off[slot][lane] = inc - len;   // lane 9: len = 0 -> total
    if(lane < 9) S.sc_base[slot][lane] = base;
    if(lane == 0){
        S.sc_flag[slot] = staged ? 0 : 2;
        if(b == clen){ S.next_q = q + (int)gridDim.x; S.next_a = 0; }
        else S.next_a = b;
    }
    __syncwarp();
    if(lane == 0){
        if(staged) bbx_mbar_arrive_expect_tx(&S.full[slot], 16u * (unsigned)total);
        else bbx_mbar_arrive(&S.full[slot]);
    }
    __syncwarp();
    if(staged && lane < 9 && len > 0) bbx_bulk_g2s(S.pool + (inc - len), pos + base, 16u * (unsigned)len, &S.full[slot]);
}

template<int SPH_EOS>
__global__ void __launch_bounds__(BBX_LT, BBX_LIST_MINB) k_cell_lists_density(StepParams P, DevGrid g, DevState *st, const int *__restrict__ occ_cells,
        const float4 *__restrict__ pos, float4 *__restrict__ vel, const int *__restrict__ cell_start,
        unsigned short *__restrict__ nbr, int *__restrict__ nbr_cnt, float *__restrict__ pressure, float4 *__restrict__ posq,
        float4 *__restrict__ rec, HaloDst H)
{
    __shared__ ListShared S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ListWarp W; W.cand = S.cand[warp]; W.spi = S.spi[warp]; W.rows = S.rows[warp]; W.tab = S.tab[warp];
    float *sovs = S.ovs[warp];
    const int n_occ = st->n_occ;
    if(threadIdx.x == 0){
        bbx_mbar_init(&S.full[0], 1); bbx_mbar_init(&S.full[1], 1);
        S.done = 0; S.next_q = blockIdx.x; S.next_a = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if(warp == 0) bbx_list_produce(S, 0, g, n_occ, occ_cells, cell_start, pos, lane);
    // half extents of a cell in the cell frame (culling of the staged candidates)
    ListFrame F;
    F.hx = 0.5f * g.lenf[0] * P.inv_h; F.hy = 0.5f * g.lenf[1] * P.inv_h; F.hz = 0.5f * g.lenf[2] * P.inv_h;
    unsigned phase = 0; // bit s = parity to wait for on full[s]
#pragma unroll 1
    for(int k = 0; ; k++){
        const int slot = k & 1;
        bbx_mbar_wait(&S.full[slot], (phase >> slot) & 1u);
        phase ^= 1u << slot;
        const int flag = S.sc_flag[slot];
        if(flag == 1) break;
        const int c = S.sc_cell[slot][warp];
        int T = 0, nc = -1, s0 = 0, m = 0;
        if(c >= 0){
            // window table: window r = (dy + 1) * 3 + (dz + 1) covers cells (cx-1..cx+1, cy+dy, cz+dz), contiguous slots
            const int cz = c / g.plane; const int rem = c - cz * g.plane; const int cy = rem / g.n[0]; const int cx = rem - cy * g.n[0];
            const int xlo = max(cx - 1, 0), xhi = min(cx + 1, g.n[0] - 1);
            F.cx = g.minf[0] + ((float)cx + 0.5f) * g.lenf[0]; F.cy = g.minf[1] + ((float)cy + 0.5f) * g.lenf[1];
            F.cz = g.minf[2] + ((float)(cz + g.zoff) + 0.5f) * g.lenf[2];
            int b = 0, len = 0;
            if(lane < 9){
                const int y = cy + lane / 3 - 1, z = cz + lane % 3 - 1;
                if(y >= 0 && y < g.n[1] && z >= 0 && z < g.n[2]){
                    const int row = y * g.n[0] + z * g.plane;
                    b = cell_start[row + xlo];
                    len = cell_start[row + xhi + 1] - b;
                }
            }
            int inc = len;
#pragma unroll
            for(int o = 1; o < 16; o <<= 1){ const int y = __shfl_up_sync(BBX_FULL, inc, o); if(lane >= o) inc += y; }
            __syncwarp();
            if(lane < 10) W.tab[lane] = inc - len;   // lane 9: len = 0 -> T
            if(lane < 9){ W.tab[10 + lane] = b; W.tab[19 + lane] = S.sc_off[slot][lane] + (b - S.sc_base[slot][lane]); }
            __syncwarp();
            T = __shfl_sync(BBX_FULL, inc, 15);
            s0 = cell_start[c]; m = cell_start[c + 1] - s0;
            if(lane == 0 && T > st->max_candidates) atomicMax(&st->max_candidates, T);
            if(flag == 0) nc = bbx_list_compact_pool(P, st, W, S.pool, T, lane, F);
        }
        // this warp is through with the pool: the last one to say so starts the copy of the next sub-chunk
        __syncwarp();
        int last = 0;
        if(lane == 0){ __threadfence_block(); last = atomicAdd(&S.done, 1) == BBX_LW - 1; }
        last = __shfl_sync(BBX_FULL, last, 0);
        if(last){
            if(lane == 0) S.done = 0;
            __syncwarp();
            bbx_list_produce(S, slot ^ 1, g, n_occ, occ_cells, cell_start, pos, lane);
        }
#pragma unroll 1
        for(int p0 = 0; p0 < m; p0 += BBX_G){
            const int mg = min(BBX_G, m - p0);
            // the group's particles: raw positions to shared memory (slow paths), cell-frame form to registers
            float4 q[BBX_G];
            {
                float4 pr = make_float4(0.f, 0.f, 0.f, 0.f), pq = make_float4(0.f, 0.f, 0.f, -1.0e30f); // unused slots accept nothing
                if(lane < mg){
                    pr = pos[s0 + p0 + lane];
                    const float ux = (pr.x - F.cx) * P.inv_h, uy = (pr.y - F.cy) * P.inv_h, uz = (pr.z - F.cz) * P.inv_h;
                    pq = make_float4(2.f * ux, 2.f * uy, 2.f * uz, 1.f - fmaf(ux, ux, fmaf(uy, uy, uz * uz)));
                }
                __syncwarp();
                if(lane < BBX_G) W.spi[lane] = pr;
#pragma unroll
                for(int ii = 0; ii < BBX_G; ii++){
                    q[ii].x = __shfl_sync(BBX_FULL, pq.x, ii); q[ii].y = __shfl_sync(BBX_FULL, pq.y, ii);
                    q[ii].z = __shfl_sync(BBX_FULL, pq.z, ii); q[ii].w = __shfl_sync(BBX_FULL, pq.w, ii);
                }
                __syncwarp();
            }
            int cnt[BBX_G]; float acc[BBX_G];
            if(bbx_list_group<false>(P, st, W, T, nc, mg, lane, F, q, pos, cnt, acc)){
                if(lane == 0) atomicAdd(&st->exact_passes, 1);
                bbx_list_group<true>(P, st, W, T, nc, mg, lane, F, q, pos, cnt, acc);
            }
            // the counts are warp-uniform: lane ii picks particle ii's
            int cn = cnt[0];
#pragma unroll
            for(int ii = 1; ii < BBX_G; ii++) cn = lane == ii ? cnt[ii] : cn;
            float sm = 0.f;
            // cap-100 slow path: lane ii redoes particle ii in the reference's order
            if(__any_sync(BBX_FULL, lane < mg && cn > BBX_MAX_NEIGHBORS)){
                if(lane < mg && cn > BBX_MAX_NEIGHBORS){
                    bbx_list_overflow_row(P, g, st, pos, cell_start, W.rows + lane * BBX_ROW, c, W.spi[lane], &cn, &sm);
                    cn = -cn; // marks "density from the slow path"
                }
                __syncwarp();
            }
            // sum the BBX_G accumulators over the 32 lanes (butterfly that halves the live values per step):
            // afterwards lane l holds the total of particle (l >> BBX_G_SHIFT) & (BBX_G - 1)
            {
                int bit = 16;
#pragma unroll
                for(int half = BBX_G / 2; half >= 1; half >>= 1){
                    const bool up = lane & bit;
#pragma unroll
                    for(int kk = 0; kk < half; kk++){
                        const float send = up ? acc[kk] : acc[kk + half], keep = up ? acc[kk + half] : acc[kk];
                        acc[kk] = keep + __shfl_xor_sync(BBX_FULL, send, bit);
                    }
                    bit >>= 1;
                }
#pragma unroll
                for(; bit >= 1; bit >>= 1) acc[0] += __shfl_xor_sync(BBX_FULL, acc[0], bit);
            }
            const int idx = (lane >> BBX_G_SHIFT) & (BBX_G - 1);
            // count / slow-path sum of particle idx live in lane idx
            const int cni = __shfl_sync(BBX_FULL, cn, idx);
            const float smi = __shfl_sync(BBX_FULL, sm, idx);
            if(!(lane & ((1 << BBX_G_SHIFT) - 1)) && idx < mg){
                const int i = s0 + p0 + idx;
                const int cabs = abs(cni); const float sum = cni < 0 ? smi : acc[0];
                nbr_cnt[i] = cabs;
                const float rho = P.mass * P.w_std_c * sum;
                // rho rides in vel.w and in the 32-byte gather record (x, y, z, rho | vx, vy, vz, -) of the viscosity
                // sweep; the grid fill wrote the record's x and v
                reinterpret_cast<float *>(vel)[4 * (size_t)i + 3] = rho;
                reinterpret_cast<float *>(rec)[8 * (size_t)i + 3] = rho;
                if(i < H.n_first || i >= H.hi_begin){
                    // boundary plane of a slab: the complete record goes to the neighbour's ghost slot as well
                    const float4 pr = W.spi[idx], vv = vel[i];
                    bbx_halo_store(H, 0, 2, i, 0, make_float4(pr.x, pr.y, pr.z, rho));
                    bbx_halo_store(H, 0, 2, i, 1, make_float4(vv.x, vv.y, vv.z, 0.f));
                }
                if(SPH_EOS){
                    // Tait EOS, ComputePressureValue (sph_equations3.cpp:7-18)
                    float p = P.eos_scale * (powf(rho / P.rho0, P.eos_exponent) - 1.f);
                    if(p < 0.f) p *= P.neg_pressure_scale;
                    pressure[i] = p;
                    const float4 pi = W.spi[idx];
                    const float qq = p / (rho * rho);
                    posq[i] = make_float4(pi.x, pi.y, pi.z, qq);
                    reinterpret_cast<float *>(rec)[8 * (size_t)i + 7] = qq; // the SPH force sweep gathers (x, rho | v, p / rho^2)
                }
            }
            // lists of the group: shared rows -> global, chunk-transposed (chunk ch of particle i is the uint4
            // ((i >> 5) * 13 + ch) * 32 + (i & 31)): consecutive lanes = consecutive particles of one chunk
            {
                const uint4 *sl = reinterpret_cast<const uint4 *>(W.rows);
                uint4 *gl = reinterpret_cast<uint4 *>(nbr);
                const int ii = lane & (BBX_G - 1);
                const int full = abs(__shfl_sync(BBX_FULL, cn, ii));
                if(ii < mg){
                    const int i = s0 + p0 + ii;
                    uint4 *dst = gl + ((size_t)(i >> 5) * BBX_NBR_CHUNKS) * 32 + (i & 31);
                    for(int ch = lane / BBX_G; ch * 8 < full; ch += 32 / BBX_G) dst[(size_t)ch * 32] = sl[ii * BBX_NBR_CHUNKS + ch];
                }
            }
            __syncwarp(); // the rows are rewritten by the next group
        }
    }
}
