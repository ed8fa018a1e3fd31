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
#include "bbx_lists.cuh"
#include "bbx_lists_tp.cuh"

#define BBX_BS 128  // threads per CTA of the particle kernels

// ------------------------------------------------------------------------- A: neighbour-grid build

// A1: cell hash per particle + histogram + jump detection.  Thread 0 also resets the per-sub-step
// statistics and the flag slots of the *next* epoch (see DevState).
// movemask[o] collects, for the particles that LEAVE old cell o, the directions they leave in (bit = 9 (dy + 1) +
// 3 (dx + 1) + (dz + 1) of new - old cell, i.e. the position of the new cell in o's neighbour list): the fill then
// walks only the old segments that really send something to a cell instead of all 27.
__global__ void __launch_bounds__(256) k_hash_count(int n_all, int n_lo, int n_own, const float4 *__restrict__ pos, const int *__restrict__ oldcell,
                                                    int *__restrict__ newcell, int *__restrict__ count,
                                                    DevGrid g, DevState *st, int have_old, int par,
                                                    unsigned long long *__restrict__ scan_status, int scan_tiles, unsigned *__restrict__ movemask)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    // slab engines: the counts of the LAST grid update are on the device (n_all < 0), the launch covers the capacity
    if(n_all < 0){ n_lo = st->n_glo; n_own = st->n_own; n_all = n_lo + n_own + st->n_ghi; }
    if(idx == 0){
        st->n_own_prev = n_own;
        st->rebuild_flag[par ^ 1] = 0; st->jump_flag[par ^ 1] = 0; st->lost[par ^ 1] = 0;
        st->overflow = 0; st->clamped = 0; st->nan_count = 0; st->max_force_bits = 0; st->max_err_bits = 0;
        st->qn[0] = 0; st->qn[1] = 0; st->qn[2] = 0; st->qn[3] = 0; st->scan_ticket = 0; st->n_occ = 0;
        st->exact_passes = 0; st->max_candidates = 0; st->unstaged_tiles = 0;
    }
    if(idx < scan_tiles) scan_status[idx] = 0ull;
    if(idx >= n_all) return;
    // slots [-n_lo, 0) hold the lower ghost plane, [0, n_own) the owned particles, [n_own, ...) the upper ghosts
    const int i = idx - n_lo;
    const bool owned = i >= 0 && i < n_own;
    float4 p = pos[i];
    int ux, uy, uz;
    int c = bbx_hash(g, p.x, p.y, p.z, &ux, &uy, &uz);
    if(c < 0){
        int gz = uz + g.zoff;
        if(ux < 0 || ux >= g.n[0] || uy < 0 || uy >= g.n[1] || gz < 0 || gz >= g.gnz){
            // outside the domain: the reference would index out of bounds here (AssertA compiled out);
            // clamp to the nearest cell and raise the sticky error flag
            ux = min(max(ux, 0), g.n[0] - 1); uy = min(max(uy, 0), g.n[1] - 1); gz = min(max(gz, 0), g.gnz - 1);
            uz = gz - g.zoff;
            if(owned) st->error = BBX_ERR_OUT_OF_DOMAIN;
        }
        if(uz < 0 || uz >= g.n[2]){
            // beyond the slab's halo: a ghost that left towards its owner's side is simply dropped; an owned
            // particle can only get here by moving two planes in one sub-step, which the halo cannot follow
            if(owned) st->error = BBX_ERR_OUT_OF_DOMAIN;
            newcell[i] = -1;
            return;
        }
        c = ux + uy * g.n[0] + uz * g.plane;
    }
    newcell[i] = c;
    atomicAdd(&count[c], 1);
    if(have_old){
        int oc = oldcell[i];
        if(oc != c){
            int oz = oc / g.plane; int rem = oc - oz * g.plane; int oy = rem / g.n[0]; int ox = rem - oy * g.n[0];
            const int dx = ux - ox, dy = uy - oy, dz = uz - oz;
            if(abs(dx) > 1 || abs(dy) > 1 || abs(dz) > 1){ if(owned){ st->jump_flag[par] = 1; atomicAdd(&st->lost[par], 1); } }
            else atomicOr(&movemask[oc], 1u << (9 * (dy + 1) + 3 * (dx + 1) + (dz + 1)));
        }
    }
}

// A2: single-pass exclusive scan of the per-cell counts (decoupled look-back over tiles of 2048 cells).
// The scanned value packs (occupied cells so far) << 31 | (particles so far), so that the same pass also
// emits the compact list of occupied cells the fill kernel iterates over.  count[] is zeroed on the way
// out (it is the histogram of the next sub-step and the cursor of the full rebuild).  count / start are
// passed offset to the first OWNED cell c0 of a slab engine (c0 = 0 for a single domain): owned slots
// start at 0, the ghost planes are laid out around them by k_ghost_table.
#define SCAN_TILE 2048
#define SCAN_FLAG_AGG (1ull << 62)
#define SCAN_FLAG_INC (2ull << 62)
#define SCAN_VALUE_MASK ((1ull << 62) - 1)
__global__ void __launch_bounds__(256) k_scan_cells(int *__restrict__ count, int total, int c0,
        unsigned long long *scan_status, DevState *st, int *__restrict__ start, int *__restrict__ occ_cells)
{
    __shared__ unsigned long long ws[8];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned s_tile;
    if(threadIdx.x == 0) s_tile = atomicAdd(&st->scan_ticket, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int base = tile * SCAN_TILE + threadIdx.x * (SCAN_TILE / 256);
    int v[SCAN_TILE / 256];
    unsigned long long s = 0;
    if(base + SCAN_TILE / 256 <= total && (((size_t)count) & 15) == 0){ // (a slab's first owned cell may be unaligned)
        int4 a = *reinterpret_cast<const int4 *>(count + base), b = *reinterpret_cast<const int4 *>(count + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        *reinterpret_cast<int4 *>(count + base) = make_int4(0, 0, 0, 0);
        *reinterpret_cast<int4 *>(count + base + 4) = make_int4(0, 0, 0, 0);
    }else{
#pragma unroll
        for(int k = 0; k < SCAN_TILE / 256; k++){ int idx = base + k; v[k] = 0; if(idx < total){ v[k] = count[idx]; count[idx] = 0; } }
    }
#pragma unroll
    for(int k = 0; k < SCAN_TILE / 256; k++) s += (unsigned long long)v[k] + (v[k] > 0 ? (1ull << 31) : 0ull);
    unsigned long long x = s;
    for(int o = 1; o < 32; o <<= 1){ unsigned long long y = __shfl_up_sync(0xffffffffu, x, o); if(lane >= o) x += y; }
    if(lane == 31) ws[warp] = x;
    __syncthreads();
    unsigned long long woff = 0, tile_sum = 0;
#pragma unroll
    for(int k = 0; k < 8; k++){ unsigned long long w = ws[k]; if(k < warp) woff += w; tile_sum += w; }
    if(warp == 0){
        // publish the tile aggregate, then look back over the predecessors (one warp, 32 tiles per probe)
        if(lane == 0){
            unsigned long long pub = (tile == 0 ? SCAN_FLAG_INC : SCAN_FLAG_AGG) | tile_sum;
            atomicExch(&scan_status[tile], pub);
        }
        unsigned long long prefix = 0;
        int look = (int)tile - 1;
        while(look >= 0){
            int idx = look - lane;
            unsigned long long stv = SCAN_FLAG_INC; // tiles before 0 count as an inclusive zero
            if(idx >= 0){
                do{ stv = *((volatile unsigned long long *)&scan_status[idx]); }while((stv >> 62) == 0);
            }
            unsigned inc = __ballot_sync(0xffffffffu, (stv >> 62) == 2);
            int first_inc = inc ? (__ffs(inc) - 1) : 32;  // nearest predecessor with an inclusive prefix
            unsigned long long val = (lane <= first_inc) ? (stv & SCAN_VALUE_MASK) : 0ull;
            for(int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
            prefix += val;
            if(inc) break;
            look -= 32;
        }
        if(lane == 0){
            if(tile != 0) atomicExch(&scan_status[tile], SCAN_FLAG_INC | (prefix + tile_sum));
            s_prefix = prefix;
        }
    }
    __syncthreads();
    unsigned long long run = s_prefix + woff + x - s;
#pragma unroll
    for(int k = 0; k < SCAN_TILE / 256; k++){
        int idx = base + k;
        if(idx < total){
            start[idx] = (int)(run & 0x7fffffffull);
            if(v[k] > 0) occ_cells[(int)(run >> 31)] = idx + c0;
            run += (unsigned long long)v[k] + (v[k] > 0 ? (1ull << 31) : 0ull);
        }
    }
    if(tile == gridDim.x - 1 && threadIdx.x == 255){
        // the last tile's last thread holds the grand total = owned particles after this update
        start[total] = (int)(run & 0x7fffffffull);
        st->n_own = (int)(run & 0x7fffffffull);
        st->n_occ = (int)(run >> 31);
        // more particles than slots (a slab that gained too many by migration): the fill / scatter / gather kernels
        // below write nothing and the host turns the sticky flag into BBX_ERR_CAPACITY
        if((int)(run & 0x7fffffffull) > st->cap) st->error = BBX_ERR_CAPACITY;
    }
}

// A3 (incremental path): 8 lanes per *occupied new* cell c walk the <= 27 old segments in the reference's
// neighbour order (y outer, x middle, z inner: Grid::GetNeighborListFor, grid.h:555-593) and append, in
// old chain order, the particles whose new cell is c (Grid::DistributeToCellOpt, grid.h:449-494) --
// ballot/popc compaction, no atomics, so the order is deterministic and equal to the reference's.
// The matching lanes move the particle payload straight into the new sorted arrays.
__global__ void __launch_bounds__(256) k_fill_incremental(DevGrid g, const DevState *st, int par, int ignore_flags,
        const int *__restrict__ occ_cells,
        const int *__restrict__ start_old, const int *__restrict__ start_new, const int *__restrict__ newcell,
        const float4 *__restrict__ pos_old, const float4 *__restrict__ vel_old, const int *__restrict__ pid_old,
        float4 *__restrict__ pos_new, float4 *__restrict__ vel_new, int *__restrict__ pid_new, int *__restrict__ cell_new,
        float4 *__restrict__ rec, const unsigned *__restrict__ movemask)
{
    // full rebuild path takes over.  (Slab engines run the fill regardless: their flags are still being reduced
    // over the ranks on a side stream; a full rebuild, if it comes, overwrites everything written here.)
    if(!ignore_flags && (st->rebuild_flag[par] | st->jump_flag[par])) return;
    if(st->n_own > st->cap) return;
    const int n_occ = st->n_occ;
    const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
    const unsigned gshift = grp * 8;
    const int groups_total = (gridDim.x * blockDim.x) >> 3;
    // all four groups of a warp iterate together (ballots are warp wide); a group without work idles
    for(int w0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 4; w0 < n_occ; w0 += groups_total){
        const int w = w0 + grp;
        const bool have = w < n_occ;
        int c = have ? occ_cells[w] : 0;
        int dst = 0, want = 0;
        if(have){ dst = start_new[c]; want = start_new[c + 1] - dst; }
        int cz = c / g.plane; int rem = c - cz * g.plane; int cy = rem / g.n[0]; int cx = rem - cy * g.n[0];
        // lane `sub` owns neighbours k = sub, sub + 8, sub + 16, sub + 24 (k < 27) in reference order
        int seg_s[4], seg_len[4];
#pragma unroll
        for(int q = 0; q < 4; q++){
            int k = sub + 8 * q;
            seg_s[q] = 0; seg_len[q] = 0;
            if(have && k < 27){
                int dy = k / 9 - 1, dx = (k / 3) % 3 - 1, dz = k % 3 - 1;
                int x = cx + dx, y = cy + dy, z = cz + dz;
                if(x >= 0 && x < g.n[0] && y >= 0 && y < g.n[1] && z >= 0 && z < g.n[2]){
                    int nb = x + y * g.n[0] + z * g.plane;
                    // neighbour k sends particles here only if some of its particles left in direction c - nb,
                    // which is entry 26 - k of ITS neighbour list (k = 13: the cell itself, the stayers)
                    if(k == 13 || ((movemask[nb] >> (26 - k)) & 1u)){
                        seg_s[q] = start_old[nb];
                        seg_len[q] = start_old[nb + 1] - seg_s[q];
                    }
                }
            }
        }
        // neighbours with something to walk, as a bit set over the reference order: per lane, OR over the group's 8
        // lanes, and the union over the warp's 4 groups drives the (warp-uniform) loop in ascending order
        unsigned act = 0;
#pragma unroll
        for(int q = 0; q < 4; q++) if(seg_len[q] > 0) act |= 1u << (sub + 8 * q);
        act |= __shfl_xor_sync(0xffffffffu, act, 1); act |= __shfl_xor_sync(0xffffffffu, act, 2); act |= __shfl_xor_sync(0xffffffffu, act, 4);
        unsigned wact = act | __shfl_xor_sync(0xffffffffu, act, 8);
        wact |= __shfl_xor_sync(0xffffffffu, wact, 16);
        int found = 0;
#pragma unroll 1
        while(wact){
            const int k = __ffs(wact) - 1; wact &= wact - 1;
            const int q = k >> 3, kk = k & 7;
            const int sq = q == 0 ? seg_s[0] : (q == 1 ? seg_s[1] : (q == 2 ? seg_s[2] : seg_s[3]));
            const int lq = q == 0 ? seg_len[0] : (q == 1 ? seg_len[1] : (q == 2 ? seg_len[2] : seg_len[3]));
            const int s = __shfl_sync(0xffffffffu, sq, kk, 8);
            int len = __shfl_sync(0xffffffffu, lq, kk, 8);
            if(found >= want) len = 0;
            // trip count = longest segment among the 4 groups
            int maxlen = len;
            maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, 8));
            maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, 16));
#pragma unroll 1
            for(int b = 0; b < maxlen; b += 8){
                int j = s + b + sub;
                bool m = (b + sub < len) && (newcell[j] == c);
                unsigned bal = (__ballot_sync(0xffffffffu, m) >> gshift) & 0xffu;
                if(m){
                    int d = dst + found + __popc(bal & ((1u << sub) - 1u));
                    const float4 pp = pos_old[j], vv = vel_old[j];
                    pos_new[d] = pp;
                    vel_new[d] = vv;
                    rec[2 * (size_t)d] = pp; rec[2 * (size_t)d + 1] = vv; // gather record (rho follows in the list build)
                    pid_new[d] = pid_old[j];
                    cell_new[d] = c;
                }
                found += __popc(bal);
            }
        }
    }
}

// A3' (full rebuild path, rare: Setup and the big-move rule): chains in ascending particle id
// (Grid::DistributeByParticle, grid.h:390-407).  scatter with atomics -> per-cell sort by id -> gather.
// Grid-stride kernels: launched with a small grid every sub-step, they return at once unless the flags
// (device side, no host round trip) ask for the rebuild.
__global__ void __launch_bounds__(256) k_full_scatter(int n_all, int n_lo, DevGrid g, const DevState *st, int par, int force, const int *__restrict__ newcell,
        const int *__restrict__ start_new, int *__restrict__ cursor, int *__restrict__ perm)
{
    if(!force && !(st->rebuild_flag[par] | st->jump_flag[par])) return;
    if(st->n_own > st->cap) return;
    if(n_all < 0){ n_lo = st->n_glo; n_all = n_lo + st->n_own_prev + st->n_ghi; } // (n_own is already the NEW count: the scan has run)
    for(int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_all; idx += gridDim.x * blockDim.x){
        int i = idx - n_lo;
        int c = newcell[i];
        if(c < g.c_own0 || c >= g.c_own1) continue; // left the slab (its new owner picks it up from its ghost plane)
        int k = atomicAdd(&cursor[c], 1);
        perm[start_new[c] + k] = i;
    }
}
// pid_old = null: order by the OLD SLOT instead of the particle id (bbx_append_particles: old chains first, in
// their order, then the appended particles in id order -- DistributeByParticleList, grid.h:358-387).  split >= 0 (append on
// a slab engine, where the appended particles were kept by an atomic cursor and sit in arbitrary slot order behind the old
// ones): slots below `split` order by slot, the others behind them by particle id.
__device__ __forceinline__ int bbx_sort_key(const int *__restrict__ pid_old, int split, int id0, int pa){
    if(split >= 0) return pa < split ? pa : split + (pid_old[pa] - id0);
    return pid_old ? pid_old[pa] : pa;
}
__global__ void __launch_bounds__(256) k_full_sort_cells(DevGrid g, const DevState *st, int par, int force, const int *__restrict__ start_new,
        const int *__restrict__ pid_old, int *__restrict__ perm, int *__restrict__ cursor, int split, int id0)
{
    if(!force && !(st->rebuild_flag[par] | st->jump_flag[par])) return;
    const bool over = st->n_own > st->cap;
    for(int c = g.c_own0 + blockIdx.x * blockDim.x + threadIdx.x; c < g.c_own1; c += gridDim.x * blockDim.x){
        cursor[c] = 0; // back to an all-zero histogram for the next sub-step
        if(over) continue;
        int s = start_new[c], e = start_new[c + 1];
        for(int a = s + 1; a < e; a++){ // insertion sort by original id (segments are a dozen long)
            int pa = perm[a]; int ka = bbx_sort_key(pid_old, split, id0, pa);
            int b = a - 1;
            while(b >= s && bbx_sort_key(pid_old, split, id0, perm[b]) > ka){ perm[b + 1] = perm[b]; b--; }
            perm[b + 1] = pa;
        }
    }
}
// bbx_append_particles after stepping: cells of the old particles are the RECORDED ones (chains are not
// re-hashed by an append), the new particles [n_old, n_old + k) hash their positions; histogram for the scan.
__global__ void __launch_bounds__(256) k_append_hash(int n_old, int k, const float4 *__restrict__ pos, const int *__restrict__ oldcell,
        int *__restrict__ newcell, int *__restrict__ count, DevGrid g, DevState *st,
        unsigned long long *__restrict__ scan_status, int scan_tiles)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i == 0){ st->scan_ticket = 0; st->n_occ = 0; }
    if(i < scan_tiles) scan_status[i] = 0ull;
    if(i >= n_old + k) return;
    int c;
    if(i < n_old) c = oldcell[i];
    else{
        float4 p = pos[i];
        int ux, uy, uz;
        c = bbx_hash(g, p.x, p.y, p.z, &ux, &uy, &uz);
        if(c < 0){ // outside the domain (the reference would index out of bounds): clamp + sticky error
            ux = min(max(ux, 0), g.n[0] - 1); uy = min(max(uy, 0), g.n[1] - 1); uz = min(max(uz, 0), g.n[2] - 1);
            c = ux + uy * g.n[0] + uz * g.plane;
            st->error = BBX_ERR_OUT_OF_DOMAIN;
        }
    }
    newcell[i] = c;
    atomicAdd(&count[c], 1);
}
__global__ void __launch_bounds__(256) k_full_gather(const DevState *st, int par, int force, const int *__restrict__ perm,
        const int *__restrict__ newcell,
        const float4 *__restrict__ pos_old, const float4 *__restrict__ vel_old, const int *__restrict__ pid_old,
        float4 *__restrict__ pos_new, float4 *__restrict__ vel_new, int *__restrict__ pid_new, int *__restrict__ cell_new,
        float4 *__restrict__ rec)
{
    if(!force && !(st->rebuild_flag[par] | st->jump_flag[par])) return;
    const int n = st->n_own;
    if(n > st->cap) return;
    for(int d = blockIdx.x * blockDim.x + threadIdx.x; d < n; d += gridDim.x * blockDim.x){
        int j = perm[d];
        const float4 pp = pos_old[j], vv = vel_old[j];
        pos_new[d] = pp;
        vel_new[d] = vv;
        rec[2 * (size_t)d] = pp; rec[2 * (size_t)d + 1] = vv;
        pid_new[d] = pid_old[j];
        cell_new[d] = newcell[j];
    }
}

// ---- slab engines: sizes of the boundary planes after the scan, and the ghost planes' part of the cell table
// With the halo push: the counts also go into the neighbours' mailboxes (my first plane's size + my owned count
// to the lower neighbour's UPPER-side mailbox, my last plane's to the upper neighbour's LOWER-side one) and
// the grid-counts flag is raised there -- no send / recv, no host round trip for the exchange of sizes.
__global__ void k_slab_counts(DevGrid g, DevState *st, const int *__restrict__ start_new, int has_lo, int has_hi,
        int *mail_lo, int *mail_hi, unsigned *flag_lo, unsigned *flag_hi, unsigned seq){
    const int n = st->n_own;
    const int nf = has_lo ? start_new[g.c_own0 + g.plane] : 0;
    const int nl = has_hi ? n - start_new[g.c_own1 - g.plane] : 0;
    st->n_first = nf; st->n_last = nl;
    if(mail_lo){ mail_lo[0] = nf; mail_lo[1] = n; }
    if(mail_hi){ mail_hi[0] = nl; mail_hi[1] = n; }
    __threadfence_system();
    if(flag_lo) *(volatile unsigned *)flag_lo = seq;
    if(flag_hi) *(volatile unsigned *)flag_hi = seq;
}
// After the neighbours' counts have arrived (k_halo_wait on the counts flag): ghost-plane sizes and the neighbours' owned
// counts from MY mailbox into DevState, where every later kernel reads them -- the host is not involved.  A ghost plane
// that would not fit (mine, or mine in the neighbour's memory) raises the sticky capacity error; the push kernel then
// writes nothing, so no allocation is overrun.
__global__ void k_slab_plan(DevState *st, const int *__restrict__ mail, int has_lo, int has_hi, int peer_gc_lo, int peer_gc_hi, int launch_bound){
    const int glo = has_lo ? mail[0] : 0, ghi = has_hi ? mail[BBX_HALO_MAIL] : 0;
    st->n_glo = glo; st->n_ghi = ghi;
    st->peer_n[0] = has_lo ? mail[1] : 0; st->peer_n[1] = has_hi ? mail[BBX_HALO_MAIL + 1] : 0;
    // (launch_bound: the per-particle kernels of this sub-step are launched over that many slots -- the host chose it from an
    // older count plus a margin; should the slab have grown past it, the sub-step is incomplete and must not go unnoticed)
    if(glo > st->gcap || ghi > st->gcap || (has_lo && st->n_first > peer_gc_lo) || (has_hi && st->n_last > peer_gc_hi) || st->n_own > st->cap
       || st->n_own > launch_bound){
        st->error = BBX_ERR_CAPACITY;
        // the run is flagged and stops at the next API call; until then every kernel of this engine becomes a no-op, so that
        // nothing indexes past an allocation or past its launch
        st->n_glo = 0; st->n_ghi = 0; st->n_own = 0; st->n_occ = 0; st->n_first = 0; st->n_last = 0;
    }
}
// host path (send / recv transport): the same fields from host values
__global__ void k_slab_plan_host(DevState *st, int glo, int ghi, int peer_lo, int peer_hi){
    st->n_glo = glo; st->n_ghi = ghi; st->peer_n[0] = peer_lo; st->peer_n[1] = peer_hi;
}
// Boundary planes of the freshly ordered arrays (cell-table slice, x, v, id) -> the neighbours' ghost slots: up to 8
// contiguous ranges whose sizes are read from DevState (k_slab_plan), 16-byte words where the alignment allows, else 4-byte.
struct PushPtrs {
    const int *tab_first, *tab_last;        // my cell-table slices over the first / last owned plane (plane + 1 entries)
    const float4 *pos, *vel; const int *pid; // my freshly ordered arrays (slot 0)
    int *gtab_lo, *gtab_hi;                  // where the slices go: lower neighbour's UPPER ghost table, upper neighbour's LOWER one
    float4 *pos_lo, *vel_lo, *pos_hi, *vel_hi; int *pid_lo, *pid_hi; // the neighbours' arrays (their slot 0)
    int has_lo, has_hi, plane;
};
__device__ __forceinline__ void bbx_push_range(const void *src, void *dst, long long b, int tid, int nth){
    const char *s = (const char *)src; char *d = (char *)dst;
    if(b <= 0) return;
    if(((((size_t)s) | ((size_t)d) | (size_t)b) & 15) == 0){
        const uint4 *s4 = (const uint4 *)s; uint4 *d4 = (uint4 *)d;
        for(long long i = tid; i < (b >> 4); i += nth) d4[i] = s4[i];
    }else{
        const int *s1 = (const int *)s; int *d1 = (int *)d;
        for(long long i = tid; i < (b >> 2); i += nth) d1[i] = s1[i];
    }
}
__global__ void __launch_bounds__(256) k_push_planes(PushPtrs Q, const DevState *st){
    if(st->error == BBX_ERR_CAPACITY) return;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int n = st->n_own, nf = st->n_first, nl = st->n_last;
    const long long tb = 4ll * (Q.plane + 1);
    if(Q.has_lo){
        const int at = st->peer_n[0]; // my first plane follows the lower neighbour's owned slots
        bbx_push_range(Q.tab_first, Q.gtab_lo, tb, tid, nth);
        bbx_push_range(Q.pos, Q.pos_lo + at, 16ll * nf, tid, nth);
        bbx_push_range(Q.vel, Q.vel_lo + at, 16ll * nf, tid, nth);
        bbx_push_range(Q.pid, Q.pid_lo + at, 4ll * nf, tid, nth);
    }
    if(Q.has_hi){
        bbx_push_range(Q.tab_last, Q.gtab_hi, tb, tid, nth);
        bbx_push_range(Q.pos + (n - nl), Q.pos_hi - nl, 16ll * nl, tid, nth);
        bbx_push_range(Q.vel + (n - nl), Q.vel_hi - nl, 16ll * nl, tid, nth);
        bbx_push_range(Q.pid + (n - nl), Q.pid_hi - nl, 4ll * nl, tid, nth);
    }
}
// recv_lo / recv_hi: the neighbour's slice of ITS cell table over the plane it sent (plane + 1 entries each).
// Lower ghost cells end at slot 0 (negative starts), upper ghost cells begin at n_own.
__global__ void __launch_bounds__(256) k_ghost_table(DevGrid g, const DevState *st, int has_lo, int has_hi,
        const int *__restrict__ recv_lo, const int *__restrict__ recv_hi,
        int *__restrict__ start_new, int *__restrict__ cell_new, int *__restrict__ count)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if(c >= g.plane) return;
    const int n_own = st->n_own;
    if(has_lo){
        int s = recv_lo[c] - recv_lo[g.plane], e = recv_lo[c + 1] - recv_lo[g.plane];
        start_new[c] = s;
        count[c] = 0;
        for(int j = s; j < e; j++) cell_new[j] = c;
    }
    if(has_hi){
        int cc = g.c_own1 + c;
        int s = n_own + recv_hi[c] - recv_hi[0], e = n_own + recv_hi[c + 1] - recv_hi[0];
        start_new[cc] = s;
        count[cc] = 0;
        for(int j = s; j < e; j++) cell_new[j] = cc;
        if(c == g.plane - 1) start_new[cc + 1] = e;
    }
}

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
__device__ __forceinline__ uint4 *bbx_chunk_ptr(unsigned short *__restrict__ nbr, int i, int chunk){
    return reinterpret_cast<uint4 *>(nbr) + ((size_t)(i >> 5) * BBX_NBR_CHUNKS + chunk) * 32 + (i & 31);
}

// ------------------------------------------------------------- list walking used by sweeps 2, 3, 4
// Each thread keeps the 9 run bases of its cell in shared memory (dynamic index by run id).
#define BBX_LIST_PROLOGUE()                                                                        \
    __shared__ int sbase[9 * BBX_BS];                                                              \
    const int n_ = bbx_count(P);                                                                   \
    const int blk_ = bbx_part_block(P, bbx_part_map(P, BBX_BS, n_), blockIdx.x);                   \
    int i = blk_ * BBX_BS + threadIdx.x;                                                           \
    bool live = blk_ >= 0 && i < n_;                                                               \
    int cnt = 0;                                                                                   \
    if(live){                                                                                      \
        int base[9], end[9];                                                                       \
        bbx_runs(g, cell_start, cell[i], base, end);                                               \
        _Pragma("unroll") for(int r = 0; r < 9; r++) sbase[r * BBX_BS + threadIdx.x] = base[r];    \
        cnt = nbr_cnt[i];                                                                          \
    }                                                                                              \
    const uint4 *lp = reinterpret_cast<const uint4 *>(nbr) + ((size_t)(i >> 5) * BBX_NBR_CHUNKS) * 32 + (i & 31);

// Full 8-entry chunks run without per-entry guards (the next chunk is requested before the current one is
// consumed); only the last, partial chunk tests k < cnt.  STRIDE = threads per CTA (column stride of sbase).
// a list chunk is read exactly once per sweep: keep it out of L1, whose lines the scattered gathers want
__device__ __forceinline__ uint4 bbx_load_chunk(const uint4 *p){
#ifndef BBX_LIST_L1ALLOC  // (measured: -2..3 % on each of the three sweeps, profiles/r02_notes.md)
    uint4 v;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
#else
    return *p;
#endif
}
template<int STRIDE, typename F>
__device__ __forceinline__ void bbx_for_each_neighbor(const uint4 *__restrict__ lp, int cnt, const int *sbase_col, F &&body){
    const int full = cnt >> 3;
    uint4 ch = make_uint4(0u, 0u, 0u, 0u);
    if(cnt > 0) ch = bbx_load_chunk(lp);
    for(int c = 0; c < full; c++){
        const uint4 cur = ch;
        if((c + 1) * 8 < cnt) ch = bbx_load_chunk(lp + (size_t)(c + 1) * 32);
        const unsigned wv[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
        for(int t = 0; t < 8; t++){
            const unsigned e = (t & 1) ? (wv[t >> 1] >> 16) : (wv[t >> 1] & 0xffffu);
            const int j = sbase_col[(e >> BBX_RUN_SHIFT) * STRIDE] + (int)(e & BBX_RUN_MASK);
            body(j);
        }
    }
    const int rest = cnt & 7;
    if(rest){
        const unsigned wv[4] = {ch.x, ch.y, ch.z, ch.w};
#pragma unroll
        for(int t = 0; t < 7; t++){
            if(t < rest){
                const unsigned e = (t & 1) ? (wv[t >> 1] >> 16) : (wv[t >> 1] & 0xffffu);
                const int j = sbase_col[(e >> BBX_RUN_SHIFT) * STRIDE] + (int)(e & BBX_RUN_MASK);
                body(j);
            }
        }
    }
}
#define BBX_LIST_FOREACH(J, ...) bbx_for_each_neighbor<BBX_BS>(lp, cnt, sbase + threadIdx.x, [&](int J) __VA_ARGS__ );

// ------------------------------------------------------ shared-memory staging of a tile's neighbourhood
// The three hot sweeps (viscosity + predict, predicted pressure, pressure force) run one CTA per TILE of
// BBX_TS consecutive slots.  Slots are sorted by cell id, so the tile covers the cell range [c_lo, c_hi] and,
// for each of the 9 (dy, dz) offsets, everything any of its particles can reference lies in the contiguous
// slot range of the cells [c_lo - 1 + off, c_hi + 1 + off] (off = dy nx + dz nx ny).  The CTA copies the
// gathered field of those 9 ranges into shared memory with coalesced 16-byte loads -- ~10 records per
// particle instead of ~49 scattered gathers through L1 -- TRANSPOSED to one array per component: the list
// walk then reads word j of each component with LDS.32, whose bank is j mod 32, and the 32 lanes of a warp
// (consecutive particles of 2-3 cells) reference slots of one narrow window, so the loads are nearly
// conflict free (float4 records read with LDS.128 were measured at 2.3 wavefronts per quarter warp: ncu,
// profiles/).  A tile whose neighbourhood does not fit (CAP slots: sparse spray next to a dense bulk) walks
// its lists against global memory instead.
#define BBX_TS 256            // threads = particles per CTA of the staged sweeps
#define BBX_STAGE_CAP 3200    // staged slots per tile (a dense tile needs ~9 x (256 + 2 x 12) = 2520)
#define BBX_STAGE_HDR 128     // bytes: tile run table
#define BBX_STAGE_BYTES(NC) (BBX_STAGE_HDR + 9 * BBX_TS * 4 + BBX_STAGE_CAP * (NC) * 4)

// Stage NC components of `src` (REC float4 per slot; component k = float k of the slot's record) for the tile
// of this CTA.  Every thread of the CTA calls this (no early exit before it).  Returns the column of
// per-thread run bases for the list walk (stride BBX_TS): an index into the component arrays
// stage[k * BBX_STAGE_CAP + j] if *staged (CTA-uniform), else a global slot index.
template<int REC, int NC>
__device__ __forceinline__ const int *bbx_stage_tile(const DevGrid &g, int n, int blk, const int *__restrict__ cell, const int *__restrict__ cell_start,
        const float4 *__restrict__ src, unsigned char *smem, DevState *st_, bool *staged, const float **stage_out)
{
    int *ttab = reinterpret_cast<int *>(smem);             // [0..9] staged offset of each run (exclusive prefix, [9] = total), [10..18] first slot
    int *sbase = reinterpret_cast<int *>(smem + BBX_STAGE_HDR);
    float *stage = reinterpret_cast<float *>(smem + BBX_STAGE_HDR + 9 * BBX_TS * 4);
    const int tid = threadIdx.x, i0 = blk * BBX_TS, i = i0 + tid;
    if(tid < 32){
        const int lane = tid;
        const int c_lo = cell[i0], c_hi = cell[min(i0 + BBX_TS, n) - 1];
        int b = 0, len = 0;
        if(lane < 9){
            const int off = (lane / 3 - 1) * g.n[0] + (lane % 3 - 1) * g.plane;
            const int ca = max(c_lo - 1 + off, 0), cb = min(c_hi + 1 + off, g.total - 1);
            if(cb >= ca){ b = cell_start[ca]; len = cell_start[cb + 1] - b; }
        }
        int inc = len;
#pragma unroll
        for(int o = 1; o < 16; o <<= 1){ int y = __shfl_up_sync(0xffffffffu, inc, o); if(lane >= o) inc += y; }
        if(lane < 10) ttab[lane] = inc - len;   // lane 9: len = 0 -> total
        if(lane < 9) ttab[10 + lane] = b;
        if(lane == 9 && inc > BBX_STAGE_CAP) atomicAdd(&st_->unstaged_tiles, 1);
    }
    int base[9], end[9];
    if(i < n) bbx_runs(g, cell_start, cell[i], base, end);
    __syncthreads();
    const bool st = ttab[9] <= BBX_STAGE_CAP;
    if(st){
        // first BBX_TS slots of every run: all loads in flight before the first store
        float4 v[9][REC];
#pragma unroll
        for(int r = 0; r < 9; r++){
            const int len = ttab[r + 1] - ttab[r];
            if(tid < len){
                const float4 *q = src + (ptrdiff_t)(ttab[10 + r] + tid) * REC;
#pragma unroll
                for(int h = 0; h < REC; h++) v[r][h] = q[h];
            }
        }
#pragma unroll
        for(int r = 0; r < 9; r++){
            const int o = ttab[r], len = ttab[r + 1] - o;
            if(tid < len){
#pragma unroll
                for(int k = 0; k < NC; k++){
                    const float4 w = v[r][k >> 2];
                    stage[k * BBX_STAGE_CAP + o + tid] = (k & 3) == 0 ? w.x : ((k & 3) == 1 ? w.y : ((k & 3) == 2 ? w.z : w.w));
                }
            }
        }
        // the rest of long runs
        for(int r = 0; r < 9; r++){
            const int o = ttab[r], len = ttab[r + 1] - o, b = ttab[10 + r];
            for(int t = tid + BBX_TS; t < len; t += BBX_TS){
                const float4 *q = src + (ptrdiff_t)(b + t) * REC;
#pragma unroll
                for(int h = 0; h < REC; h++){
                    const float4 w = q[h];
                    if(4 * h + 0 < NC) stage[(4 * h + 0) * BBX_STAGE_CAP + o + t] = w.x;
                    if(4 * h + 1 < NC) stage[(4 * h + 1) * BBX_STAGE_CAP + o + t] = w.y;
                    if(4 * h + 2 < NC) stage[(4 * h + 2) * BBX_STAGE_CAP + o + t] = w.z;
                    if(4 * h + 3 < NC) stage[(4 * h + 3) * BBX_STAGE_CAP + o + t] = w.w;
                }
            }
        }
    }
    if(i < n){
#pragma unroll
        for(int r = 0; r < 9; r++) sbase[r * BBX_TS + tid] = st ? base[r] - ttab[10 + r] + ttab[r] : base[r];
    }
    __syncthreads();
    *staged = st; *stage_out = stage;
    return sbase + tid;
}

// max over the sub-step of a non-negative float (its bits order like unsigned): warp max first, then one
// atomic per warp; a partially active warp falls back to one atomic per thread
__device__ __forceinline__ void bbx_atomic_max_warp(unsigned *addr, float v){
    if(__activemask() == 0xffffffffu){
        for(int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if((threadIdx.x & 31) == 0) atomicMax(addr, __float_as_uint(v));
    }else atomicMax(addr, __float_as_uint(v));
}

// d and 1/d of a squared distance without the IEEE sqrt sequence: MUFU.RSQ + 1 multiply (2 ulp)
// (raw MUFU through PTX: rsqrtf / __fdividef without fast-math carry denormal fix-up code, ~6 instructions)
__device__ __forceinline__ float bbx_rsqrt_approx(float x){ float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float bbx_rcp_approx(float x){ float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float bbx_rsqrt_safe(float d2){ return bbx_rsqrt_approx(fmaxf(d2, 1.0e-30f)); }

// --------------------------------- C+D: non-pressure forces + first prediction (sweep 2)
// f_i = m g - c_drag v_i + mu m^2 sum_j (v_j - v_i) d2W_spiky(d) / rho_j   (ComputeNonPressureForceFor,
// sph_equations3.cpp:80-110), then x* = x + dt (v + dt/m f), collide (restitution 0)
// (PredictVelocityAndPositionFor with is_first, pcisph_equations3.cpp:3-28).  A particle the FP32
// pre-check cannot clear of every collider is queued for k_collide_predict.
// (One thread per particle, neighbours gathered straight from global memory: the shared-memory staging that
// k_pressure uses was measured SLOWER here -- 7 words per neighbour cost ~19 bank-conflicted wavefronts from
// shared memory against one LDG.E.256 through L1; profiles/r01_notes.md.)
__global__ void __launch_bounds__(BBX_BS) k_force_np_predict(StepParams P, DevGrid g, DevState *st, const DevCullSet *__restrict__ cull,
        const float4 *__restrict__ pos, const float4 *__restrict__ vel, const float4 *__restrict__ rec, const int *__restrict__ cell,
        const int *__restrict__ cell_start, const unsigned short *__restrict__ nbr, const int *__restrict__ nbr_cnt,
        float4 *__restrict__ force, float4 *__restrict__ pred, int *__restrict__ queue, HaloDst H)
{
    BBX_LIST_PROLOGUE();
    if(!live) return;
    float4 pi = pos[i]; float4 vi = vel[i];
    float ax = 0.f, ay = 0.f, az = 0.f;
    BBX_LIST_FOREACH(j, {
        // x_j, rho_j, v_j in ONE 256-bit gather (LDG.E.256) from the 32-byte records the list build wrote
        float4 pj, vj;
        asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=f"(pj.x), "=f"(pj.y), "=f"(pj.z), "=f"(vj.w), "=f"(vj.x), "=f"(vj.y), "=f"(vj.z), "=f"(pj.w) : "l"(rec + 2 * (ptrdiff_t)j));
        float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        float d = d2 * bbx_rsqrt_safe(d2);
        float x = fmaxf(0.f, fmaf(-d, P.inv_h, 1.f));
        float w = x * bbx_rcp_approx(vj.w);
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
    pred[i] = make_float4(tpx, tpy, tpz, 0.f);
    if(!bbx_cull(*cull, tpx, tpy, tpz, P.radius)) queue[atomicAdd(&st->qn[P.part == 1 ? 2 : 0], 1)] = i; // (k_collide_predict pushes its halo copy)
    else{ H = bbx_halo_resolve(H, st); bbx_halo_store(H, 0, 1, i, 0, make_float4(tpx, tpy, tpz, 0.f)); }
}

// x* of one particle from its forces + the exact collider response (restitution 0)
__device__ __forceinline__ float4 bbx_predict_exact(const StepParams &P, const DevColliderSet &cs, float4 pi, float4 vi, float fx, float fy, float fz){
    float k = P.dt * P.inv_mass;
    float tvx = vi.x + k * fx, tvy = vi.y + k * fy, tvz = vi.z + k * fz;
    float tpx = pi.x + P.dt * tvx, tpy = pi.y + P.dt * tvy, tpz = pi.z + P.dt * tvz;
    bbx_resolve_collision(cs, (double)P.radius, 0.0, &tpx, &tpy, &tpz, &tvx, &tvy, &tvz);
    return make_float4(tpx, tpy, tpz, 0.f);
}
// queued particles of k_force_np_predict: redo the prediction with the exact FP64 response
__global__ void __launch_bounds__(128) k_collide_predict(StepParams P, const DevState *st, const DevColliderSet *__restrict__ cs,
        const int *__restrict__ queue, const float4 *__restrict__ pos, const float4 *__restrict__ vel,
        const float4 *__restrict__ force, float4 *__restrict__ pred, HaloDst H)
{
    const int qn = st->qn[P.part == 1 ? 2 : 0];
    H = bbx_halo_resolve(H, st);
    for(int q = blockIdx.x * blockDim.x + threadIdx.x; q < qn; q += gridDim.x * blockDim.x){
        int i = queue[q];
        float4 f = force[i];
        const float4 x = bbx_predict_exact(P, *cs, pos[i], vel[i], f.x, f.y, f.z);
        pred[i] = x;
        bbx_halo_store(H, 0, 1, i, 0, x);
    }
}

// later iterations of the predict-correct loop ("correct" mode): x* from f_np + f_p, no neighbour sum
__global__ void __launch_bounds__(256) k_predict_again(StepParams P, const DevColliderSet *__restrict__ cs, const DevCullSet *__restrict__ cull,
        const float4 *__restrict__ pos, const float4 *__restrict__ vel, const float4 *__restrict__ force,
        const float4 *__restrict__ force_p, float4 *__restrict__ pred)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= bbx_count(P)) return;
    float4 pi = pos[i], vi = vel[i], f = force[i], fp = force_p[i];
    float fx = f.x + fp.x, fy = f.y + fp.y, fz = f.z + fp.z;
    float k = P.dt * P.inv_mass;
    float tvx = vi.x + k * fx, tvy = vi.y + k * fy, tvz = vi.z + k * fz;
    float tpx = pi.x + P.dt * tvx, tpy = pi.y + P.dt * tvy, tpz = pi.z + P.dt * tvz;
    if(bbx_cull(*cull, tpx, tpy, tpz, P.radius)) pred[i] = make_float4(tpx, tpy, tpz, 0.f);
    else pred[i] = bbx_predict_exact(P, *cs, pi, vi, fx, fy, fz);
}

// ----------------------------------------------------------- E: predicted density -> pressure (sweep 3)
// rho*_i = m sum_j W_std(|x*_i - x*_j|) over the same list; p += delta (rho* - rho0), negative increments
// scaled by negativePressureScale (PredictPressureFor, pcisph_equations3.cpp:60-90).
// Writes posq = (x_i, p_i / rho*_i^2) for the pressure-force sweep.
extern __shared__ __align__(128) unsigned char bbx_dyn_smem[];

__global__ void __launch_bounds__(BBX_TS, 4) k_pressure(StepParams P, DevGrid g, DevState *st, int first,
        const float4 *__restrict__ pos, const float4 *__restrict__ pred, const int *__restrict__ cell,
        const int *__restrict__ cell_start, const unsigned short *__restrict__ nbr, const int *__restrict__ nbr_cnt,
        float *__restrict__ pressure, float *__restrict__ rho_pred, float *__restrict__ rho_err, float4 *__restrict__ posq, HaloDst H)
{
    bool staged; const float *S;
    const int n = bbx_count(P);
    const int blk = bbx_part_block(P, bbx_part_map(P, BBX_TS, n), blockIdx.x);
    if(blk < 0) return; // (the launch covers the capacity of a slab engine / the largest boundary pass)
    const int *scol = bbx_stage_tile<1, 3>(g, n, blk, cell, cell_start, pred, bbx_dyn_smem, st, &staged, &S);
    const int i = blk * BBX_TS + threadIdx.x;
    if(i >= n) return;
    const int cnt = nbr_cnt[i];
    const uint4 *lp = reinterpret_cast<const uint4 *>(nbr) + ((size_t)(i >> 5) * BBX_NBR_CHUNKS) * 32 + (i & 31);
    const float4 pi = pred[i];
    float sum = 0.f;
    auto pair = [&](const float4 pj){
        float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        float x = fmaxf(0.f, fmaf(-d2, P.inv_h2, 1.f));
        sum = fmaf(x * x, x, sum);
    };
    if(staged) bbx_for_each_neighbor<BBX_TS>(lp, cnt, scol, [&](int j){ const float *q = S + j; pair(make_float4(q[0], q[BBX_STAGE_CAP], q[2 * BBX_STAGE_CAP], 0.f)); });
    else bbx_for_each_neighbor<BBX_TS>(lp, cnt, scol, [&](int j){ pair(pred[j]); });
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
    const float4 xq = make_float4(x0.x, x0.y, x0.z, (rho2 < 1e-8f) ? __int_as_float(0x7fc00000) : p / rho2);
    posq[i] = xq;
    H = bbx_halo_resolve(H, st); // (late: the plane sizes are only needed here, not across the neighbour loop)
    bbx_halo_store(H, 0, 1, i, 0, xq);
    bbx_atomic_max_warp(&st->max_err_bits, fabsf(err));
}

// integration of one particle: v += dt f/m, x += dt v, collide (restitution), domain clamp
// (TimeIntegrationFor, sph_equations3.cpp:283-328).  EXACT = 0 is the FP32 fast path, valid when the
// pre-checks cleared the colliders and the domain faces.
template<int EXACT>
__device__ __forceinline__ void bbx_integrate_one(const StepParams &P, const DevGrid &g, DevState *st, const DevColliderSet *cs,
        float4 pi, float4 v, float fx, float fy, float fz, float4 *pos_out, float4 *vel_out)
{
    float vx = v.x + P.dt * (fx * P.inv_mass), vy = v.y + P.dt * (fy * P.inv_mass), vz = v.z + P.dt * (fz * P.inv_mass);
    float px = pi.x + P.dt * vx, py = pi.y + P.dt * vy, pz = pi.z + P.dt * vz;
    if(EXACT){
        bbx_resolve_collision(*cs, (double)P.radius, (double)P.restitution, &px, &py, &pz, &vx, &vy, &vz);
        // domain clamp (sph_equations3.cpp:317-326)
        if(!inside_bounds(v3(px, py, pz), v3(g.min[0], g.min[1], g.min[2]), v3(g.max[0], g.max[1], g.max[2]))){
            double r = (double)P.radius;
            px = (float)clampd(px, g.min[0] + r, g.max[0] - r);
            py = (float)clampd(py, g.min[1] + r, g.max[1] - r);
            pz = (float)clampd(pz, g.min[2] + r, g.max[2] - r);
            atomicAdd(&st->clamped, 1);
        }
    }
    *pos_out = make_float4(px, py, pz, 0.f);
    *vel_out = make_float4(vx, vy, vz, v.w);
}
// big-move rule (sph_equations3.cpp:330-335) and the non-finite counter, on the final position
__device__ __forceinline__ void bbx_integrate_flags(const StepParams &P, DevState *st, float4 pi, float4 po){
    float mx = po.x - pi.x, my = po.y - pi.y, mz = po.z - pi.z;
    if(sqrtf(mx * mx + my * my + mz * mz) >= P.min_cell_len09) st->rebuild_flag[P.par ^ 1] = 1;
    if(!(isfinite(po.x) && isfinite(po.y) && isfinite(po.z))) atomicAdd(&st->nan_count, 1);
}
// max |f| of the sub-step for the CFL scan (particle.h:591-597)
__device__ __forceinline__ void bbx_reduce_max_force(DevState *st, float f2){
    bbx_atomic_max_warp(&st->max_force_bits, sqrtf(f2));
}

// -------------------------------------- F+G: pressure force (+ accumulate, integrate, collide) (sweep 4)
// f_p,i = - m^2 sum_{j != i} (p_i/rho*_i^2 + p_j/rho*_j^2) gradW_spiky  with current positions
// (PredictPressureForceFor, pcisph_equations3.cpp:114-155); INTEGRATE: f += f_p, v += dt f/m, x += dt v,
// collide (restitution 0.6), domain clamp, big-move flag (AccumulateForcesFor + TimeIntegrationFor,
// pcisph_equations3.cpp:179-197, sph_equations3.cpp:283-339).  Positions are updated in place: neighbours
// are read from posq, never from pos, so there is no read/write race.  Particles near a collider or a
// domain face (FP32 pre-check undecided) are queued for k_collide_integrate.
template<int INTEGRATE>
__global__ void __launch_bounds__(BBX_BS) k_pressure_force(StepParams P, DevGrid g, DevState *st, const DevCullSet *__restrict__ cull,
        float4 *__restrict__ pos, float4 *__restrict__ vel, const float4 *__restrict__ posq, const int *__restrict__ cell,
        const int *__restrict__ cell_start, const unsigned short *__restrict__ nbr, const int *__restrict__ nbr_cnt,
        float4 *__restrict__ force, float4 *__restrict__ force_p, int *__restrict__ queue, HaloDst H)
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
        // j == i and coincident points (d ~ 0) and NaN-marked neighbours contribute nothing
        float inv_d = (d2 > 1.0e-16f && pj.w == pj.w) ? bbx_rsqrt_approx(d2) : 0.f;
        float x = fmaxf(0.f, fmaf(-d2 * inv_d, P.inv_h, 1.f));
        float w = (qi + pj.w) * (x * x) * inv_d;
        w = (inv_d > 0.f) ? w : 0.f;
        tx = fmaf(dx, w, tx); ty = fmaf(dy, w, ty); tz = fmaf(dz, w, tz);
    })
    float s = -P.mass2 * P.dw_spiky_c;
    float fpx = s * tx, fpy = s * ty, fpz = s * tz;
    force_p[i] = make_float4(fpx, fpy, fpz, 0.f);
    if(INTEGRATE){
        float4 f = force[i]; float4 v = vel[i];
        float fx = f.x + fpx, fy = f.y + fpy, fz = f.z + fpz;
        force[i] = make_float4(fx, fy, fz, 0.f);
        bbx_reduce_max_force(st, fmaf(fx, fx, fmaf(fy, fy, fz * fz)));
        float4 po, vo;
        bbx_integrate_one<0>(P, g, st, nullptr, pi, v, fx, fy, fz, &po, &vo);
        if(bbx_cull(*cull, po.x, po.y, po.z, P.radius) && bbx_inside_domain_certain(*cull, po.x, po.y, po.z)){
            bbx_integrate_flags(P, st, pi, po);
            pos[i] = po; vel[i] = vo;
            H = bbx_halo_resolve(H, st);
            bbx_halo_store(H, 0, 1, i, 0, po); bbx_halo_store(H, 1, 1, i, 0, vo);
        }else{
            queue[atomicAdd(&st->qn[P.part == 1 ? 3 : 1], 1)] = i; // pos / vel stay untouched: the exact kernel redoes the update
        }
    }
}
// queued particles of k_pressure_force<1>: the exact FP64 collider response + domain clamp
__global__ void __launch_bounds__(128) k_collide_integrate(StepParams P, DevGrid g, DevState *st, const DevColliderSet *__restrict__ cs,
        const int *__restrict__ queue, float4 *__restrict__ pos, float4 *__restrict__ vel, const float4 *__restrict__ force, HaloDst H)
{
    const int qn = st->qn[P.part == 1 ? 3 : 1];
    H = bbx_halo_resolve(H, st);
    for(int q = blockIdx.x * blockDim.x + threadIdx.x; q < qn; q += gridDim.x * blockDim.x){
        int i = queue[q];
        float4 f = force[i];
        float4 po, vo;
        float4 pi = pos[i];
        bbx_integrate_one<1>(P, g, st, cs, pi, vel[i], f.x, f.y, f.z, &po, &vo);
        bbx_integrate_flags(P, st, pi, po);
        pos[i] = po; vel[i] = vo;
        bbx_halo_store(H, 0, 1, i, 0, po); bbx_halo_store(H, 1, 1, i, 0, vo);
    }
}

// integrate alone ("correct" mode after the loop, and the SPH step)
__global__ void __launch_bounds__(256) k_integrate(StepParams P, DevGrid g, DevState *st, const DevColliderSet *__restrict__ cs, const DevCullSet *__restrict__ cull,
        float4 *__restrict__ pos, float4 *__restrict__ vel, float4 *__restrict__ force, const float4 *__restrict__ force_p)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= bbx_count(P)) return;
    float4 pi = pos[i], v = vel[i], f = force[i];
    float fx = f.x, fy = f.y, fz = f.z;
    if(force_p){ float4 fp = force_p[i]; fx += fp.x; fy += fp.y; fz += fp.z; force[i] = make_float4(fx, fy, fz, 0.f); }
    bbx_reduce_max_force(st, fmaf(fx, fx, fmaf(fy, fy, fz * fz)));
    float4 po, vo;
    bbx_integrate_one<0>(P, g, st, nullptr, pi, v, fx, fy, fz, &po, &vo);
    if(!(bbx_cull(*cull, po.x, po.y, po.z, P.radius) && bbx_inside_domain_certain(*cull, po.x, po.y, po.z)))
        bbx_integrate_one<1>(P, g, st, cs, pi, v, fx, fy, fz, &po, &vo);
    bbx_integrate_flags(P, st, pi, po);
    pos[i] = po; vel[i] = vo;
}

// ------------------------------------------------------------------ SPH (non-PCI) force sweep
// ComputeAllForcesFor (sph_equations3.cpp:184-272) with Jacobi semantics: gravity + drag + viscosity
// (only inside the spiky support and j != i) + pressure force from rho, p of this sub-step.
__global__ void __launch_bounds__(BBX_BS) k_sph_forces(StepParams P, DevGrid g,
        const float4 *__restrict__ posq, const float4 *__restrict__ vel, const float4 *__restrict__ rec, const int *__restrict__ cell,
        const int *__restrict__ cell_start, const unsigned short *__restrict__ nbr, const int *__restrict__ nbr_cnt,
        float4 *__restrict__ force)
{
    BBX_LIST_PROLOGUE();
    if(!live) return;
    float4 pi = posq[i]; float4 vi = vel[i];
    float tx = 0.f, ty = 0.f, tz = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
    float qi = pi.w;
    BBX_LIST_FOREACH(j, {
        // x_j, rho_j, v_j, p_j / rho_j^2 in ONE 256-bit gather from the records the list build completed
        float4 pj, vj;
        asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=f"(pj.x), "=f"(pj.y), "=f"(pj.z), "=f"(pj.w), "=f"(vj.x), "=f"(vj.y), "=f"(vj.z), "=f"(vj.w) : "l"(rec + 2 * (ptrdiff_t)j));
        float dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
        float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        float inv_d = (d2 > 1.0e-16f) ? bbx_rsqrt_approx(d2) : 0.f;
        float x = (j != i) ? fmaxf(0.f, fmaf(-d2 * inv_d, P.inv_h, 1.f)) : 0.f;
        float w = (qi + vj.w) * (x * x) * inv_d;
        tx = fmaf(dx, w, tx); ty = fmaf(dy, w, ty); tz = fmaf(dz, w, tz);
        float wv = x * bbx_rcp_approx(pj.w);
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
    if(i >= bbx_count(P)) return;
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
// slab engines: keep only the particles whose cell plane this slab owns (any superset of them may be passed,
// e.g. the whole scene); ids = global particle ids (null: the index in the arrays).  The slot order is
// arbitrary -- the full rebuild that follows orders every cell by ascending id.
__global__ void __launch_bounds__(256) k_upload_slab(int n, const void *__restrict__ pos, const void *__restrict__ vel, const int *__restrict__ ids,
        int is_f64, DevGrid g, int cap, DevState *st, float4 *__restrict__ dpos, float4 *__restrict__ dvel, int *__restrict__ pid)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    float p[3], v[3];
    for(int k = 0; k < 3; k++){
        if(is_f64){ p[k] = (float)((const double *)pos)[3 * (size_t)i + k]; v[k] = (float)((const double *)vel)[3 * (size_t)i + k]; }
        else{ p[k] = ((const float *)pos)[3 * (size_t)i + k]; v[k] = ((const float *)vel)[3 * (size_t)i + k]; }
    }
    int ux, uy, uz;
    bbx_hash(g, p[0], p[1], p[2], &ux, &uy, &uz);
    int gz = min(max(uz + g.zoff, 0), g.gnz - 1) - g.zoff; // out-of-domain z is clamped like k_hash_count does
    if(gz < g.own_z0 || gz >= g.own_z1) return;
    int slot = atomicAdd(&st->n_own, 1);
    if(slot >= cap){ st->error = BBX_ERR_CAPACITY; return; }
    dpos[slot] = make_float4(p[0], p[1], p[2], 0.f);
    dvel[slot] = make_float4(v[0], v[1], v[2], 0.f);
    pid[slot] = ids ? ids[i] : i;
}
// overwrite pos/vel of existing particles: slot i holds particle pid[i]
__global__ void __launch_bounds__(256) k_overwrite(int n, const int *__restrict__ pid, const void *__restrict__ pos, const void *__restrict__ vel,
                                                   int is_f64, float4 *__restrict__ dpos, float4 *__restrict__ dvel)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    size_t id = pid ? (size_t)pid[i] : (size_t)i; // pid = null: rows in slot order (bbx_overwrite_owned)
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
// (pid = null: keep the slot order -- bbx_download_owned)
__global__ void __launch_bounds__(256) k_download(int n, const int *__restrict__ pid, const float4 *__restrict__ src4,
                                                  const float *__restrict__ src1, const int *__restrict__ srci, int comps, int is_f64, void *__restrict__ dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    size_t id = pid ? (size_t)pid[i] : (size_t)i;
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
// positions AND velocities in one pass (bbx_download_state): row = particle id, or the slot when pid = null
__global__ void __launch_bounds__(256) k_download_state(int n, const int *__restrict__ pid, const float4 *__restrict__ pos, const float4 *__restrict__ vel,
                                                        int is_f64, void *__restrict__ dpos, void *__restrict__ dvel)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    size_t id = pid ? (size_t)pid[i] : (size_t)i;
    const float4 p = pos[i], v = vel[i];
    if(is_f64){
        double *a = (double *)dpos + 3 * id, *b = (double *)dvel + 3 * id;
        a[0] = p.x; a[1] = p.y; a[2] = p.z; b[0] = v.x; b[1] = v.y; b[2] = v.z;
    }else{
        float *a = (float *)dpos + 3 * id, *b = (float *)dvel + 3 * id;
        a[0] = p.x; a[1] = p.y; a[2] = p.z; b[0] = v.x; b[1] = v.y; b[2] = v.z;
    }
}
// ContinuousParticleSetBuilder3::MapGridEmit's per-cell test (src/core/grid.h:1367-1407) for a batch of template points:
// size of the CURRENT chain of the point's cell, and whether a particle of that chain lies closer than d --
// Distance(pj, pi) = sqrt(|pj - pi|^2) in FP64 like the reference (geometry.h:632-633, 797-800) on the FP32-stored positions.
// cells are GLOBAL cell ids; a cell this (slab) engine does not own answers size -1.
__global__ void __launch_bounds__(256) k_query_cells(int n, DevGrid g, const int *__restrict__ cells, const double *__restrict__ points, double d,
        const int *__restrict__ cell_start, const float4 *__restrict__ pos, int *__restrict__ cell_size, int *__restrict__ blocked)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= n) return;
    const int c = cells[k] - g.zoff * g.plane; // local cell id
    if(cells[k] < 0 || c < g.c_own0 || c >= g.c_own1){ cell_size[k] = -1; blocked[k] = 0; return; }
    const int s = cell_start[c], e = cell_start[c + 1];
    const double px = points[3 * (size_t)k], py = points[3 * (size_t)k + 1], pz = points[3 * (size_t)k + 2];
    int hit = 0;
    for(int j = s; j < e && !hit; j++){
        const float4 q = pos[j];
        const double x = __dsub_rn((double)q.x, px), y = __dsub_rn((double)q.y, py), z = __dsub_rn((double)q.z, pz);
        const double l2 = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
        if(__dsqrt_rn(l2) < d) hit = 1;
    }
    cell_size[k] = e - s; blocked[k] = hit;
}
// Shape::ClosestDistance of one collider at a batch of points (parity tests of the mesh BVH: bbx_collider_distance)
__global__ void __launch_bounds__(128) k_collider_distance(int n, const DevColliderSet *__restrict__ cs, int index, const double *__restrict__ points, double *__restrict__ out){
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= n) return;
    out[k] = closest_distance(cs->c[index], v3(points[3 * (size_t)k], points[3 * (size_t)k + 1], points[3 * (size_t)k + 2]));
}
__global__ void __launch_bounds__(256) k_export_cells(int total, const int *__restrict__ cell_start, int *__restrict__ cell_count){
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if(c < total) cell_count[c] = cell_start[c + 1] - cell_start[c];
}
// Decode the *stored* compact lists into the reference's bucket order (neighbour cells y/x/z, chain
// order inside a cell) with original ids: what Bucket::pids holds after UpdateParticlesBuckets.
__global__ void __launch_bounds__(BBX_BS) k_export_neighbors(int n, DevGrid g, const int *__restrict__ pid, const int *__restrict__ cell,
        const int *__restrict__ cell_start, const unsigned short *__restrict__ nbr, const int *__restrict__ nbr_cnt,
        int *__restrict__ counts, int *__restrict__ ids, int compact)
{
    int i = blockIdx.x * BBX_BS + threadIdx.x;
    if(i >= n) return;
    int c = cell[i];
    int base[9], end[9];
    bbx_runs(g, cell_start, c, base, end);
    int cz = c / g.plane; int rem = c - cz * g.plane; int cy = rem / g.n[0]; int cx = rem - cy * g.n[0];
    (void)cz; (void)cy;
    int cnt = nbr_cnt[i];
    size_t id = compact ? (size_t)i : (size_t)pid[i]; // output row: slot order (slab engines) or particle id
    int *out = ids + id * BBX_MAX_NEIGHBORS;
    // key = (reference rank of the neighbour cell) << 40 | slot: sort ascending (insertion, <= 100 items)
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
        float4 *__restrict__ pos_new, float4 *__restrict__ vel_new, int *__restrict__ pid_new, int *__restrict__ cell_new,
        float4 *__restrict__ rec)
{
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if(d >= n) return;
    int id = order[d]; int j = slot_of[id];
    const float4 pp = pos_old[j], vv = vel_old[j];
    pos_new[d] = pp; vel_new[d] = vv; rec[2 * (size_t)d] = pp; rec[2 * (size_t)d + 1] = vv;
    pid_new[d] = id; cell_new[d] = cell_of_slot_new[d];
}

// bbx_rebalance: recorded cell ids moved to another local plane numbering (dst may be src)
__global__ void __launch_bounds__(256) k_cells_shift(int n, const int *src, int *dst, int delta){
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) dst[i] = src[i] + delta;
}

// bbx_rebalance: the slots [0, n) are in cell order already (planes arrive in chain order and are laid out in plane order):
// rebuild the owned part of the cell table from the recorded cells, and the gather records of the force sweeps
__global__ void __launch_bounds__(256) k_table_from_sorted(int n, DevGrid g, const int *__restrict__ cell, int *__restrict__ start,
        const float4 *__restrict__ pos, const float4 *__restrict__ vel, float4 *__restrict__ rec)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(n == 0){ for(int c = g.c_own0 + i; c <= g.c_own1; c += gridDim.x * blockDim.x) start[c] = 0; return; }
    if(i >= n) return;
    const int c = cell[i], prev = i > 0 ? cell[i - 1] : g.c_own0 - 1;
    for(int cc = prev + 1; cc <= c; cc++) start[cc] = i;
    if(i == n - 1) for(int cc = c + 1; cc <= g.c_own1; cc++) start[cc] = n;
    rec[2 * (size_t)i] = pos[i]; rec[2 * (size_t)i + 1] = vel[i];
}
