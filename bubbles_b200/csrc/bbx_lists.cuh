// bbx list build (phase B, engine v4): per-particle neighbour lists + density, ONE WARP PER OCCUPIED CELL,
// "transposed": the 32 lanes hold 32 consecutive CANDIDATES of the cell's 27-cell neighbourhood, the loop
// runs over the cell's own particles (broadcast from shared memory).
//
// Why (ncu, profiles/r01_v2_*): with one thread per particle the candidate walk diverges (19 of 32 lanes
// active) and each lane turns its accept bits into list entries one by one.  Here
//   * every pair test of a round is one warp-wide instruction stream without divergence,
//   * the accept decisions of (round, particle) are ONE ballot; an accepted lane's list position is the
//     particle's running count + popc(ballot & lanes below) -- no atomics, deterministic flat order,
//   * candidates are read straight from global memory (coalesced inside a run), one round ahead of use;
//     the only shared-memory traffic is the broadcast of the own particle and the 2-byte list stores.
// The pair loop is ~23 instructions per (round, particle): LDS of the particle, 6 FP32 for d^2, compare,
// ballot, guard-band flag, 7 for the list position / store / count, 4 FP32 for the clamped W_std.
// (An "expand the ballots afterwards, one lane per particle" variant was measured and is 3x slower in the
// expansion than these 7 instructions: 12 of 32 lanes active, ~17 instructions per set bit.)
//
// Candidates are visited in the flat order run-major / slot order, the same order v2 produced, and lists
// keep v2's format: entry = (run << 12) | offset-in-run, chunk-transposed per 32 particles (the sweeps'
// coalesced 512 B loads).  Any neighbourhood size works (no tile to overflow).
//
// Exactness: a candidate is accepted when d2 < thr_hi (FP32).  If any accepted candidate of a pass lay
// inside the guard band [thr_lo, thr_hi] around h^2 - 1e-8, the pass is repeated with the FP64 predicate of
// the reference (IsWithinStd, kernel.cpp:229-234) deciding the band members -- ~2 % of the passes.
// Reference: Grid::DistributeParticleBucket grid.h:422-447, Bucket::Insert particle.h:44-50 (cap 100),
// ComputeDensityFor + ComputePressureValue sph_equations3.cpp:7-58.
#pragma once
#include "bbx_device.cuh"

#define BBX_LW 4                  // warps per CTA
#define BBX_LT (BBX_LW * 32)
#define BBX_BP 16                 // own particles per pass (density accumulators live in registers)
#define BBX_ROW 104               // u16 entries per list row in shared memory (13 chunks of 8)
#define BBX_FULL 0xffffffffu

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
    float4 *spi;             // [BBX_BP] own particles of the pass (unused slots: far away)
    unsigned short *rows;    // [BBX_BP][BBX_ROW] lists being built
    int *tab;                // [0..9] exclusive prefix of the 9 run lengths (tab[9] = T), [10..18] first slot of each run
    int *cnt;                // [BBX_BP]
};

// One pass: all candidates of the cell against up to 16 own particles.  Returns (warp-uniform) whether a
// provisionally accepted candidate lay inside the guard band.
template<bool EXACT>
__device__ __forceinline__ bool bbx_list_pass(const StepParams &P, DevState *st, const ListWarp &W, int T, int mp, int lane,
        const float4 *__restrict__ pos, float *acc)
{
    bool band = false;
    const unsigned lt = (1u << lane) - 1u;
    int cnt[BBX_BP];
#pragma unroll
    for(int ii = 0; ii < BBX_BP; ii++){ acc[ii] = 0.f; cnt[ii] = 0; }
    // candidate of (round 0, this lane), fetched one round ahead of its use
    int f = lane;
    float4 pj_next; unsigned short e_next;
    {
        const int fc = min(f, T - 1);
        int r = 0;
#pragma unroll
        for(int k = 1; k < 9; k++) r += (fc >= W.tab[k]) ? 1 : 0;
        const int off = fc - W.tab[r];
        pj_next = pos[W.tab[10 + r] + off];
        e_next = (unsigned short)((r << BBX_RUN_SHIFT) | min(off, BBX_MAX_RUN_LEN - 1));
        if(off >= BBX_MAX_RUN_LEN) st->error = BBX_ERR_CAPACITY;
        if(f >= T) pj_next.x = 1.0e15f;
    }
    {
#pragma unroll 1
        for(int f0 = 0; f0 < T; f0 += 32){
            const float4 pj = pj_next;
            const unsigned short entry = e_next;
            // prefetch the next round's candidate
            f += 32;
            if(f - lane < T){
                const int fc = min(f, T - 1);
                int r = 0;
#pragma unroll
                for(int k = 1; k < 9; k++) r += (fc >= W.tab[k]) ? 1 : 0;
                const int off = fc - W.tab[r];
                pj_next = pos[W.tab[10 + r] + off];
                e_next = (unsigned short)((r << BBX_RUN_SHIFT) | min(off, BBX_MAX_RUN_LEN - 1));
                if(off >= BBX_MAX_RUN_LEN) st->error = BBX_ERR_CAPACITY;
                if(f >= T) pj_next.x = 1.0e15f;
            }
#pragma unroll
            for(int g4 = 0; g4 < BBX_BP / 4; g4++){
                if(g4 * 4 < mp){
#pragma unroll
                    for(int q = 0; q < 4; q++){
                        const int ii = g4 * 4 + q;
                        const float4 pi = W.spi[ii];
                        const float dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
                        const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        bool in = d2 < P.thr_hi;
                        if(EXACT){
                            if(in && d2 >= P.thr_lo) in = bbx_within_std_exact(pi, pj, P.h2_d);
                        }else{
                            band |= in && (d2 >= P.thr_lo);
                        }
                        const unsigned msk = __ballot_sync(BBX_FULL, in);
                        // list position = entries so far + accepted lanes below this one (flat order); rows have
                        // 104 slots, the count keeps running so that the cap-100 slow path can be detected
                        const int k = min(cnt[ii] + __popc(msk & lt), BBX_ROW - 1);
                        if(in) W.rows[ii * BBX_ROW + k] = entry;
                        cnt[ii] += __popc(msk);
                        const float x = fmaxf(0.f, fmaf(-d2, P.inv_h2, 1.f));
                        acc[ii] = fmaf(x * x, x, acc[ii]);
                    }
                }
            }
        }
    }
#pragma unroll
    for(int ii = 0; ii < BBX_BP; ii++) if(lane == ii) W.cnt[ii] = cnt[ii];
    __syncwarp();
    return __any_sync(BBX_FULL, band);
}

template<int SPH_EOS>
__global__ void __launch_bounds__(BBX_LT) k_cell_lists_density(StepParams P, DevGrid g, DevState *st, const int *__restrict__ occ_cells,
        const float4 *__restrict__ pos, float4 *__restrict__ vel, const int *__restrict__ cell_start,
        unsigned short *__restrict__ nbr, int *__restrict__ nbr_cnt, float *__restrict__ pressure, float4 *__restrict__ posq,
        float4 *__restrict__ rec)
{
    __shared__ float4 s_pi[BBX_LW][BBX_BP];
    __shared__ __align__(16) unsigned short s_rows[BBX_LW][BBX_BP * BBX_ROW];
    __shared__ int s_tab[BBX_LW][20];
    __shared__ int s_cnt[BBX_LW][BBX_BP];
    __shared__ float s_ovs[BBX_LW][BBX_BP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ListWarp W; W.spi = s_pi[warp]; W.rows = s_rows[warp]; W.tab = s_tab[warp]; W.cnt = s_cnt[warp];
    float *sovs = s_ovs[warp];
    const int n_occ = st->n_occ;
    const int nwarps = gridDim.x * BBX_LW;
#pragma unroll 1
    for(int w = blockIdx.x * BBX_LW + warp; w < n_occ; w += nwarps){
        const int c = occ_cells[w];
        // run table: run r = (dy + 1) * 3 + (dz + 1) covers cells (cx-1..cx+1, cy+dy, cz+dz), contiguous slots
        int T;
        {
            int cz = c / g.plane; int rem = c - cz * g.plane; int cy = rem / g.n[0]; int cx = rem - cy * g.n[0];
            int xlo = max(cx - 1, 0), xhi = min(cx + 1, g.n[0] - 1);
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
#pragma unroll 1
        for(int p0 = 0; p0 < m; p0 += BBX_BP){
            const int mp = min(BBX_BP, m - p0);
            __syncwarp();
            if(lane < BBX_BP) W.spi[lane] = lane < mp ? pos[s0 + p0 + lane] : make_float4(-1.0e15f, -1.0e15f, -1.0e15f, 0.f);
            __syncwarp();
            float acc[BBX_BP];
            if(bbx_list_pass<false>(P, st, W, T, mp, lane, pos, acc)){
                if(lane == 0) atomicAdd(&st->exact_passes, 1);
                bbx_list_pass<true>(P, st, W, T, mp, lane, pos, acc);
            }
            // cap-100 slow path: lane ii redoes particle ii in the reference's order
            if(lane < mp){
                int cn = W.cnt[lane]; float sm = 0.f;
                if(cn > BBX_MAX_NEIGHBORS){
                    bbx_list_overflow_row(P, g, st, pos, cell_start, W.rows + lane * BBX_ROW, c, W.spi[lane], &cn, &sm);
                    cn = -cn; // marks "density from the slow path"
                }
                W.cnt[lane] = cn; sovs[lane] = sm;
            }
            // sum the 16 accumulators over the 32 lanes (butterfly that halves the live values per step):
            // afterwards lane l holds the total of particle (l >> 1) & 15
#pragma unroll
            for(int k = 0; k < 8; k++){
                const bool up = lane & 16;
                const float send = up ? acc[k] : acc[k + 8], keep = up ? acc[k + 8] : acc[k];
                acc[k] = keep + __shfl_xor_sync(BBX_FULL, send, 16);
            }
#pragma unroll
            for(int k = 0; k < 4; k++){
                const bool up = lane & 8;
                const float send = up ? acc[k] : acc[k + 4], keep = up ? acc[k + 4] : acc[k];
                acc[k] = keep + __shfl_xor_sync(BBX_FULL, send, 8);
            }
#pragma unroll
            for(int k = 0; k < 2; k++){
                const bool up = lane & 4;
                const float send = up ? acc[k] : acc[k + 2], keep = up ? acc[k + 2] : acc[k];
                acc[k] = keep + __shfl_xor_sync(BBX_FULL, send, 4);
            }
            {
                const bool up = lane & 2;
                const float send = up ? acc[0] : acc[1], keep = up ? acc[1] : acc[0];
                acc[0] = keep + __shfl_xor_sync(BBX_FULL, send, 2);
            }
            acc[0] += __shfl_xor_sync(BBX_FULL, acc[0], 1);
            __syncwarp();
            const int idx = (lane >> 1) & 15;
            if(!(lane & 1) && idx < mp){
                const int i = s0 + p0 + idx;
                int cn = W.cnt[idx]; float sum = acc[0];
                if(cn < 0){ cn = -cn; sum = sovs[idx]; }
                nbr_cnt[i] = cn;
                const float rho = P.mass * P.w_std_c * sum;
                // rho rides in vel.w and in the 32-byte gather record (x, y, z, rho | vx, vy, vz, -) of the viscosity
                // sweep; the grid fill wrote the record's x and v
                reinterpret_cast<float *>(vel)[4 * (size_t)i + 3] = rho;
                reinterpret_cast<float *>(rec)[8 * (size_t)i + 3] = rho;
                if(SPH_EOS){
                    // Tait EOS, ComputePressureValue (sph_equations3.cpp:7-18)
                    float p = P.eos_scale * (powf(rho / P.rho0, P.eos_exponent) - 1.f);
                    if(p < 0.f) p *= P.neg_pressure_scale;
                    pressure[i] = p;
                    const float4 pi = W.spi[idx];
                    posq[i] = make_float4(pi.x, pi.y, pi.z, p / (rho * rho));
                }
            }
            // lists of the pass: shared rows -> global, chunk-transposed (chunk ch of particle i is the uint4
            // ((i >> 5) * 13 + ch) * 32 + (i & 31)): consecutive lanes = consecutive particles of one chunk
            const uint4 *sl = reinterpret_cast<const uint4 *>(W.rows);
            uint4 *gl = reinterpret_cast<uint4 *>(nbr);
            for(int t = lane; t < mp * BBX_NBR_CHUNKS; t += 32){
                const int ch = t / mp, ii = t - ch * mp;
                if(ch * 8 < abs(W.cnt[ii])){
                    const int i = s0 + p0 + ii;
                    gl[((size_t)(i >> 5) * BBX_NBR_CHUNKS + ch) * 32 + (i & 31)] = sl[ii * BBX_NBR_CHUNKS + ch];
                }
            }
        }
    }
}
