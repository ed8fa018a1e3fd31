// bbx list build (phase B): per-particle neighbour lists + density, ONE WARP PER OCCUPIED CELL,
// "transposed": the 32 lanes hold 32 consecutive CANDIDATES of the cell's 27-cell neighbourhood, the own
// particles of the cell are held in REGISTERS, 8 at a time.
//
// History (ncu, profiles/): one thread per particle diverges on the candidate walk (v2); v4/v5 made the lanes
// candidates and looped over 16 own particles broadcast from shared memory -- every pair test then started
// with an LDS.128 whose latency the 4-5 resident warps per scheduler could not hide (issue slots 62 % busy,
// the rest waiting on the load).  Here
//   * the cell's candidates are staged ONCE in shared memory in the cell frame: (u_j, list entry) with
//     u = (x - cell centre) / h -- one coalesced global load and ~40 instructions per candidate per cell;
//   * a group of 8 own particles lives in registers as (2 u_i, 1 - |u_i|^2); the round loop reads one
//     candidate per lane (conflict-free LDS.128, fetched one round ahead) and runs 8 pair tests on registers;
//   * the pair test is x = 1 - d^2 / h^2 = (1 - |u_i|^2) - |u_j|^2 + 2 u_j . u_i : 3 FFMA + 1 FADD; accepted when
//     x > xacc, and the same x is the density weight (W_std ~ x^3);
//   * the accept decisions of (round, particle) are ONE ballot; an accepted lane's list position is the
//     particle's running count + popc(ballot & lanes below) -- no atomics, deterministic flat order.
// ~17 instructions per (round, particle).
//
// Candidates are visited in the flat order run-major / slot order and lists keep the format of v2:
// entry = (run << 12) | offset-in-run, chunk-transposed per 32 particles (the sweeps' coalesced 512 B loads).
// Neighbourhoods larger than BBX_CMAX candidates are staged in several chunks per group.
//
// Exactness: in the cell frame |u| <= ~2.6, so x carries an absolute error of a few 1e-7.  A candidate is
// accepted when x > xacc (d^2 certainly below h^2 - 1e-8 + band).  If any accepted candidate of a group lay
// inside the guard band (x < xband), the group is redone with the FP64 predicate of the reference
// (IsWithinStd, kernel.cpp:229-234) deciding the band members -- ~1 % of the groups.
// Reference: Grid::DistributeParticleBucket grid.h:422-447, Bucket::Insert particle.h:44-50 (cap 100),
// ComputeDensityFor + ComputePressureValue sph_equations3.cpp:7-58.
#pragma once
#include "bbx_device.cuh"

#define BBX_LW 4                  // warps per CTA
#define BBX_LT (BBX_LW * 32)
#ifndef BBX_G
#define BBX_G 4                   // own particles per group (registers): 4 or 8
#endif
#define BBX_G_SHIFT (BBX_G == 8 ? 2 : 3) // lane >> shift = particle whose density total the lane holds after the butterfly
#ifndef BBX_CMAX
#define BBX_CMAX 512              // staged candidates per chunk (multiple of 32)
#endif
#define BBX_ROW 104               // u16 entries per list row in shared memory (13 chunks of 8)
#define BBX_FULL 0xffffffffu
#ifndef BBX_LIST_MINB
#define BBX_LIST_MINB 5           // resident CTAs per SM the list kernel is compiled for
#endif

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
    unsigned short *ent;     // [BBX_CMAX] list entries of the raw candidates (first pass of the staging)
    float4 *spi;             // [BBX_G] own particles of the group, raw positions (FP64 re-check, cap-100 slow path)
    unsigned short *rows;    // [BBX_G][BBX_ROW] lists being built
    int *tab;                // [0..9] exclusive prefix of the 9 run lengths (tab[9] = T), [10..18] first slot of each run
    int *cnt;                // [BBX_G]
};

// Stage candidates of the cell's neighbourhood in flat order (run-major, slot order), starting at flat index
// f0: up to BBX_CMAX of them per call (f0 advances; f0 = T when all are consumed); returns the number staged
// (the array is padded to a multiple of 32 with candidates that are never accepted).
//   pass 1: the raw positions go global -> shared with cp.async (LDGSTS, 16 B per candidate, every copy of the
//           chunk in flight at once -- nothing waits on a load), the list entry (run, offset) beside them;
//   pass 2: in place, round by round: cell frame u = (x - centre) / h (exact subtraction in FP32: both are
//           multiples of the finer ulp and the difference is small), and a candidate farther than h from the
//           cell's box -- it cannot be a neighbour of any particle of the cell -- is dropped (ballot compaction
//           keeps the flat order): ~24 % of the 27-cell neighbourhood.
__device__ __forceinline__ int bbx_list_stage(const StepParams &P, DevState *st, const ListWarp &W, int T, int lane, int &f0,
        float cx, float cy, float cz, float hx, float hy, float hz, const float4 *__restrict__ pos)
{
    __syncwarp();
    const unsigned lt = lanemask_lt();
    const int c0 = f0, nc = min(BBX_CMAX, T - c0);
    {
        int r = 0;
        const unsigned cand_addr = (unsigned)__cvta_generic_to_shared(W.cand);
#pragma unroll 2
        for(int k = lane; k < nc; k += 32){
            const int f = c0 + k;
            while(f >= W.tab[r + 1]) r++;           // runs only move forward (tab[9] = T > f)
            const int off = f - W.tab[r];
            const float4 *src = pos + W.tab[10 + r] + off;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(cand_addr + 16u * (unsigned)k), "l"(src) : "memory");
            W.ent[k] = (unsigned short)(((unsigned)r << BBX_RUN_SHIFT) | (unsigned)min(off, BBX_MAX_RUN_LEN - 1));
            if(off >= BBX_MAX_RUN_LEN) st->error = BBX_ERR_CAPACITY;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    int n = 0;
#pragma unroll 1
    for(int k0 = 0; k0 < nc; k0 += 32){
        const int k = k0 + lane;
        const float4 raw = W.cand[min(k, nc - 1)];
        const unsigned entry = W.ent[min(k, nc - 1)];
        __syncwarp(); // every lane has read its candidate before the compacted ones overwrite this range
        const float ux = (raw.x - cx) * P.inv_h, uy = (raw.y - cy) * P.inv_h, uz = (raw.z - cz) * P.inv_h;
        const float gx = fmaxf(fabsf(ux) - hx, 0.f), gy = fmaxf(fabsf(uy) - hy, 0.f), gz = fmaxf(fabsf(uz) - hz, 0.f);
        const bool keep = (k < nc) && (fmaf(gx, gx, fmaf(gy, gy, gz * gz)) < 1.001f);
        const unsigned msk = __ballot_sync(BBX_FULL, keep);
        if(keep) W.cand[n + __popc(msk & lt)] = make_float4(ux, uy, uz, __uint_as_float(entry));
        n += __popc(msk);
    }
    f0 = c0 + nc;
    __syncwarp();
    if(n + lane < ((n + 31) & ~31)) W.cand[n + lane] = make_float4(1.0e15f, 0.f, 0.f, 0.f);
    __syncwarp();
    return n;
}

// All staged candidates against the group's (up to 8) own particles in registers.  Accumulates cnt / acc over
// chunks; xmin = smallest x among this lane's accepted pairs.
template<bool EXACT>
__device__ __forceinline__ void bbx_list_rounds(const StepParams &P, const ListWarp &W, int nc, int mg, int lane,
        const float4 (&q)[BBX_G], const float4 *__restrict__ pos, int (&cnt)[BBX_G], float (&acc)[BBX_G], float &xmin)
{
    const unsigned lt = lanemask_lt();
    // shared-memory byte address of this warp's list rows, pinned in a register (the compiler otherwise
    // rebuilds it from special registers in front of every store)
    unsigned rows_addr;
    asm volatile("mov.u32 %0, %1;" : "=r"(rows_addr) : "r"((unsigned)__cvta_generic_to_shared(W.rows)));
    float4 cn = W.cand[lane]; // (nc = 0: never used)
#pragma unroll 1
    for(int f0 = 0; f0 < nc; f0 += 32){
        const float4 cj = cn;
        if(f0 + 32 < nc) cn = W.cand[f0 + 32 + lane];
        const float a = fmaf(cj.x, cj.x, fmaf(cj.y, cj.y, cj.z * cj.z));
        const unsigned short entry = (unsigned short)__float_as_uint(cj.w);
#pragma unroll
        for(int h = 0; h < BBX_G / 4; h++){
            if(h * 4 < mg){
#pragma unroll
                for(int t = 0; t < 4; t++){
                    const int ii = h * 4 + t;
                    const float x = fmaf(cj.x, q[ii].x, fmaf(cj.y, q[ii].y, fmaf(cj.z, q[ii].z, q[ii].w))) - a;
                    bool in = x > P.xacc;
                    if(EXACT){
                        if(in && x < P.xband){
                            const unsigned e = entry;
                            in = bbx_within_std_exact(W.spi[ii], pos[W.tab[10 + (e >> BBX_RUN_SHIFT)] + (int)(e & BBX_RUN_MASK)], P.h2_d);
                        }
                    }else{
                        xmin = in ? fminf(xmin, x) : xmin;
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
            }
        }
    }
}

// One group: every chunk of the neighbourhood against the group's particles (nc >= 0: the whole neighbourhood
// is staged already, nc candidates).  Returns (warp-uniform) whether a provisionally accepted candidate lay
// inside the guard band.
template<bool EXACT>
__device__ __forceinline__ bool bbx_list_group(const StepParams &P, DevState *st, const ListWarp &W, int T, int nc, int mg, int lane,
        float cx, float cy, float cz, float hx, float hy, float hz, const float4 (&q)[BBX_G], const float4 *__restrict__ pos,
        int (&cnt)[BBX_G], float (&acc)[BBX_G])
{
#pragma unroll
    for(int ii = 0; ii < BBX_G; ii++){ acc[ii] = 0.f; cnt[ii] = 0; }
    float xmin = 1.0e30f;
    if(nc >= 0){
        bbx_list_rounds<EXACT>(P, W, nc, mg, lane, q, pos, cnt, acc, xmin);
    }else{
        int f0 = 0;
#pragma unroll 1
        while(f0 < T){
            const int n = bbx_list_stage(P, st, W, T, lane, f0, cx, cy, cz, hx, hy, hz, pos);
            bbx_list_rounds<EXACT>(P, W, n, mg, lane, q, pos, cnt, acc, xmin);
        }
    }
    return __any_sync(BBX_FULL, xmin < P.xband);
}

template<int SPH_EOS>
__global__ void __launch_bounds__(BBX_LT, BBX_LIST_MINB) k_cell_lists_density(StepParams P, DevGrid g, DevState *st, const int *__restrict__ occ_cells,
        const float4 *__restrict__ pos, float4 *__restrict__ vel, const int *__restrict__ cell_start,
        unsigned short *__restrict__ nbr, int *__restrict__ nbr_cnt, float *__restrict__ pressure, float4 *__restrict__ posq,
        float4 *__restrict__ rec, HaloDst H)
{
    __shared__ float4 s_cand[BBX_LW][BBX_CMAX + 32];
    __shared__ unsigned short s_ent[BBX_LW][BBX_CMAX];
    __shared__ float4 s_pi[BBX_LW][BBX_G];
    __shared__ __align__(16) unsigned short s_rows[BBX_LW][BBX_G * BBX_ROW];
    __shared__ int s_tab[BBX_LW][20];
    __shared__ int s_cnt[BBX_LW][BBX_G];
    __shared__ float s_ovs[BBX_LW][BBX_G];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ListWarp W; W.cand = s_cand[warp]; W.ent = s_ent[warp]; W.spi = s_pi[warp]; W.rows = s_rows[warp]; W.tab = s_tab[warp]; W.cnt = s_cnt[warp];
    float *sovs = s_ovs[warp];
    const int n_occ = st->n_occ;
    const int nwarps = gridDim.x * BBX_LW;
    H = bbx_halo_resolve(H, st);
#pragma unroll 1
    for(int w = blockIdx.x * BBX_LW + warp; w < n_occ; w += nwarps){
        const int c = occ_cells[w];
        // run table: run r = (dy + 1) * 3 + (dz + 1) covers cells (cx-1..cx+1, cy+dy, cz+dz), contiguous slots
        int T;
        float ccx, ccy, ccz; // cell centre: origin of the cell frame of the pair test
        {
            int cz = c / g.plane; int rem = c - cz * g.plane; int cy = rem / g.n[0]; int cx = rem - cy * g.n[0];
            int xlo = max(cx - 1, 0), xhi = min(cx + 1, g.n[0] - 1);
            ccx = g.minf[0] + ((float)cx + 0.5f) * g.lenf[0]; ccy = g.minf[1] + ((float)cy + 0.5f) * g.lenf[1];
            ccz = g.minf[2] + ((float)(cz + g.zoff) + 0.5f) * g.lenf[2];
            int b = 0, len = 0;
            if(lane < 9){
                int y = cy + lane / 3 - 1, z = cz + lane % 3 - 1;
                if(y >= 0 && y < g.n[1] && z >= 0 && z < g.n[2]){
                    int row = y * g.n[0] + z * g.plane;
                    b = cell_start[row + xlo];
                    len = cell_start[row + xhi + 1] - b;
                }
            }
            int inc = len;
#pragma unroll
            for(int o = 1; o < 16; o <<= 1){ int y = __shfl_up_sync(BBX_FULL, inc, o); if(lane >= o) inc += y; }
            __syncwarp();
            if(lane < 10) W.tab[lane] = inc - len;   // lane 9: len = 0 -> T
            if(lane < 9) W.tab[10 + lane] = b;
            __syncwarp();
            T = __shfl_sync(BBX_FULL, inc, 15);
        }
        const int s0 = cell_start[c], m = cell_start[c + 1] - s0;
        if(lane == 0 && T > st->max_candidates) atomicMax(&st->max_candidates, T);
        // half extents of the cell in the cell frame (culling of the staged candidates)
        const float hx = 0.5f * g.lenf[0] * P.inv_h, hy = 0.5f * g.lenf[1] * P.inv_h, hz = 0.5f * g.lenf[2] * P.inv_h;
        int nc = -1; // >= 0: the whole neighbourhood fits one stage, done once per cell
        if(T <= BBX_CMAX){ int f0 = 0; nc = bbx_list_stage(P, st, W, T, lane, f0, ccx, ccy, ccz, hx, hy, hz, pos); }
#pragma unroll 1
        for(int p0 = 0; p0 < m; p0 += BBX_G){
            const int mg = min(BBX_G, m - p0);
            // the group's particles: raw positions to shared memory (slow paths), cell-frame form to registers
            float4 q[BBX_G];
            {
                float4 pr = make_float4(0.f, 0.f, 0.f, 0.f), pq = make_float4(0.f, 0.f, 0.f, -1.0e30f); // unused slots accept nothing
                if(lane < mg){
                    pr = pos[s0 + p0 + lane];
                    const float ux = (pr.x - ccx) * P.inv_h, uy = (pr.y - ccy) * P.inv_h, uz = (pr.z - ccz) * P.inv_h;
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
            if(bbx_list_group<false>(P, st, W, T, nc, mg, lane, ccx, ccy, ccz, hx, hy, hz, q, pos, cnt, acc)){
                if(lane == 0) atomicAdd(&st->exact_passes, 1);
                bbx_list_group<true>(P, st, W, T, nc, mg, lane, ccx, ccy, ccz, hx, hy, hz, q, pos, cnt, acc);
            }
#pragma unroll
            for(int ii = 0; ii < BBX_G; ii++) if(lane == ii) W.cnt[ii] = cnt[ii];
            __syncwarp();
            // cap-100 slow path: lane ii redoes particle ii in the reference's order
            if(lane < mg){
                int cn = W.cnt[lane]; float sm = 0.f;
                if(cn > BBX_MAX_NEIGHBORS){
                    bbx_list_overflow_row(P, g, st, pos, cell_start, W.rows + lane * BBX_ROW, c, W.spi[lane], &cn, &sm);
                    cn = -cn; // marks "density from the slow path"
                }
                W.cnt[lane] = cn; sovs[lane] = sm;
            }
            // sum the BBX_G accumulators over the 32 lanes (butterfly that halves the live values per step):
            // afterwards lane l holds the total of particle (l >> BBX_G_SHIFT) & (BBX_G - 1)
            {
                int bit = 16;
#pragma unroll
                for(int half = BBX_G / 2; half >= 1; half >>= 1){
                    const bool up = lane & bit;
#pragma unroll
                    for(int k = 0; k < half; k++){
                        const float send = up ? acc[k] : acc[k + half], keep = up ? acc[k + half] : acc[k];
                        acc[k] = keep + __shfl_xor_sync(BBX_FULL, send, bit);
                    }
                    bit >>= 1;
                }
#pragma unroll
                for(; bit >= 1; bit >>= 1) acc[0] += __shfl_xor_sync(BBX_FULL, acc[0], bit);
            }
            __syncwarp();
            const int idx = (lane >> BBX_G_SHIFT) & (BBX_G - 1);
            if(!(lane & ((1 << BBX_G_SHIFT) - 1)) && idx < mg){
                const int i = s0 + p0 + idx;
                int cn = W.cnt[idx]; float sum = acc[0];
                if(cn < 0){ cn = -cn; sum = sovs[idx]; }
                nbr_cnt[i] = cn;
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
                    const float q = p / (rho * rho);
                    posq[i] = make_float4(pi.x, pi.y, pi.z, q);
                    reinterpret_cast<float *>(rec)[8 * (size_t)i + 7] = q; // the SPH force sweep gathers (x, rho | v, p / rho^2)
                }
            }
            // lists of the group: shared rows -> global, chunk-transposed (chunk ch of particle i is the uint4
            // ((i >> 5) * 13 + ch) * 32 + (i & 31)): consecutive lanes = consecutive particles of one chunk
            const uint4 *sl = reinterpret_cast<const uint4 *>(W.rows);
            uint4 *gl = reinterpret_cast<uint4 *>(nbr);
            {
                const int ii = lane & (BBX_G - 1);
                if(ii < mg){
                    const int i = s0 + p0 + ii, full = abs(W.cnt[ii]);
                    uint4 *dst = gl + ((size_t)(i >> 5) * BBX_NBR_CHUNKS) * 32 + (i & 31);
                    for(int ch = lane / BBX_G; ch * 8 < full; ch += 32 / BBX_G) dst[(size_t)ch * 32] = sl[ii * BBX_NBR_CHUNKS + ch];
                }
            }
        }
    }
}
