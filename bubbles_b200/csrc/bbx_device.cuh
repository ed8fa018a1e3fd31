// Device-side building blocks shared by the bbx kernels: grid hashing (FP64, bit-exact with the
// reference), the neighbour acceptance predicate (FP32 with an FP64 re-check inside a guard band),
// SPH kernel weights and the collider response.
//
// Reference semantics restated here (paths relative to the Bubbles tree):
//   Grid::GetHashedPosition / ExtremeEpsilon / LinearIndex   src/core/grid.h:259-298, 182-193
//   IsWithinStd / IsWithinSpiky / kernel weights            src/core/kernel.cpp:123-234
//   ColliderSet3::ResolveCollision -> CollisionHandle       src/core/collider.cpp:238-264, 3-44
//   Box / sphere / SDF closest point                        src/shapes/box.cpp:74-172, src/shapes/sphere.cpp:47-70,
//                                                           src/shapes/bvh.cpp:666-707, src/core/grid.h:1093-1169
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/bbx.h"

#define BBX_NBR_CHUNKS 13   // 13 x 8 = 104 >= 100 list entries per particle
#define BBX_RUN_SHIFT 12    // list entry = (run << 12) | offset-in-run
#define BBX_RUN_MASK 0xfffu
#define BBX_MAX_RUN_LEN 4096

// The grid an engine works on is the LOCAL grid of its z-slab: whole cell planes [zoff, zoff + n[2]) of
// the global grid (z is the slowest index of LinearIndex, so a slab is a contiguous range of global
// cell ids and local id = global id - zoff * plane).  Owned planes are local [own_z0, own_z1); a slab
// with a lower / upper neighbour carries one ghost plane on that side.  Single domain: zoff = 0,
// n[2] = gnz, everything owned.  Geometry (min, max, len) is always the global grid's, so hashing is
// bit-identical to the single-domain engine.
struct DevGrid {
    double min[3], max[3], len[3];
    float minf[3], maxf[3], lenf[3];
    int n[3];           // local cell counts (n[2] = local planes incl. ghost planes)
    int total;          // local cells
    int plane;          // n[0] * n[1]
    int zoff;           // global z of local plane 0
    int gnz;            // global n[2]
    int own_z0, own_z1; // owned local planes
    int c_own0, c_own1; // owned local cells [own_z0 * plane, own_z1 * plane)
};

// one node of the BVH over a mesh collider's triangles (built on the host, bbx_set_colliders): a leaf holds `count` > 0
// triangles starting at `first` of the reordered triangle list, an inner node its two children
struct DevBvhNode { double lo[3], hi[3]; int left, right, first, count; };

struct DevCollider {
    int type, reverse, active, pad;
    double o2w[16];
    double w2o[16];
    double size[3];
    double radius;
    double friction;
    double linvel[3];
    double angvel[3];
    int sdf_res[3];
    int pad2;
    double sdf_spacing[3];
    double sdf_origin[3];
    const double *sdf_field; // device pointer
    const float *sdf_field32; // FP32 shadow of the field (pre-check only)
    double lipschitz;         // upper bound of |grad| of the trilinear field
    // triangle mesh (BBX_COLLIDER_MESH): vertices, triangles in BVH leaf order, the BVH, bounds of the triangles
    const double *mesh_points; const int *mesh_tris; const DevBvhNode *bvh;
    int n_tris, n_nodes;
    double mesh_lo[3], mesh_hi[3];
};

struct DevColliderSet {
    int count;
    int pad;
    DevCollider c[BBX_MAX_COLLIDERS];
};

// FP32 shadow of the collider set, used by the conservative "certainly no collision" pre-check of the
// neighbour sweeps (bbx_cull).  Particles that fail the pre-check are queued and finished by the FP64
// restatement of ColliderSet3::ResolveCollision (bbx_resolve_collision) in a follow-up kernel.
struct DevCullCollider {
    int type, reverse, active, identity; // identity: object space == world space
    float w2o[12];      // rows 0..2 of WorldToObject (affine)
    float half[3];      // box half sizes
    float radius;       // sphere
    float lipschitz;    // SDF: upper bound of |grad f| of the trilinear field
    int sdf_res[3];
    float sdf_inv_spacing[3];
    float sdf_origin[3];
    const float *sdf_field32;
};
struct DevCullSet {
    int count;
    float margin;       // absolute slack covering the FP32 evaluation error
    float dom_lo[3], dom_hi[3]; // domain bounds shrunk by margin (certainly-inside test of the domain clamp)
    DevCullCollider c[BBX_MAX_COLLIDERS];
};

// Per-engine mutable device state (flags + statistics), one instance in global memory.
// Flags that one sub-step writes and the next one reads are double-buffered by the parity of the grid
// update ("epoch") that consumes them, so no kernel has to reset a flag another kernel of the same
// sub-step may be setting.
struct DevState {
    int rebuild_flag[2]; // SphParticleSet3::requiresHigherLevelUpdate: written by integrate for the next epoch
    int jump_flag[2];    // some particle moved >= 2 cells since the last grid update (chains would lose it)
    int lost[2];         // particles that jumped >= 2 cells
    int overflow;        // particles whose neighbour list hit the 100 cap
    int clamped;         // particles pushed back into the domain
    int nan_count;
    int error;           // sticky: bbx_status-like device-detected error (out of domain, run too long)
    unsigned max_force_bits; // float bits of max |f| (non-negative floats order like unsigned)
    unsigned max_err_bits;   // float bits of max |rho* - rho0|
    int n_occ;           // occupied cells found by the last scan
    int qn[4];           // collider slow-path queue lengths: [0] predict, [1] integrate; [2], [3]: the same for the boundary passes of a slab engine (their own queue)
    unsigned scan_ticket;
    int iterations;
    int n_own;           // owned particles after the last grid update / upload (slab engines)
    int n_first, n_last; // particles in the first / last owned plane (what the slab neighbours hold as ghosts)
    int exact_passes;    // list-build passes repeated with the FP64 predicate (a candidate inside the guard band)
    int max_candidates;  // largest 27-cell neighbourhood seen in the last list build
    int unstaged_tiles;  // sweep tiles (all three sweeps) whose neighbourhood exceeded the shared-memory stage
    int cap;             // particle capacity of the engine (owned slots): the fill / scatter kernels refuse to write past it
    // slab engines: what the kernels need to know about the ghost planes, kept ON THE DEVICE so that no sub-step waits for
    // the host (k_slab_plan writes them from the neighbours' mailboxes; the host reads them back lazily)
    int n_glo, n_ghi;    // particles in the lower / upper ghost plane
    int peer_n[2];       // owned particles of the lower / upper neighbour (where my boundary planes start in ITS slot space)
    int gcap;            // ghost slots per side of this engine
    int n_own_prev;      // owned particles BEFORE the running grid update (the scan overwrites n_own; the full rebuild walks the old slots)
};

// Scalars of one sub-step, passed by value to every kernel.  bbx_count(P): owned particles -- a kernel argument on
// single-domain engines, a device read on slab engines (no host round trip in their grid update).
struct StepParams {
    int n;              // launch bound of the per-particle kernels (>= owned particles; the kernels read the count from dyn)
    int n_owned;
    const DevState *dyn; // slab engines: their DevState, where the owned count lives (dyn->n_own); null: n is exact
    float h, h2, inv_h, inv_h2;
    float thr2;         // h^2 - 1e-8: acceptance threshold of IsWithinStd on d^2
    float band;         // guard band around thr2 inside which the predicate is re-evaluated in FP64
    double h_d, h2_d;   // FP64 copies for the exact re-check
    float mass, mass2, inv_mass;
    float rho0;
    float w_std_c;      // 315 / (64 pi h^3)
    float d2w_spiky_c;  // 90 / (pi h^5)
    float dw_spiky_c;   // 45 / (pi h^4)   (gradW = +c (1-d/h)^2 dir)
    float w_spiky_c;    // 15 / (pi h^3)
    float viscosity, drag;
    float gx, gy, gz;
    float dt;
    float delta;
    float neg_pressure_scale;
    float eos_scale, eos_exponent; // rho0 c^2 / gamma, gamma
    float radius;       // particle radius = spacing
    float restitution;
    float min_cell_len09; // 0.9 * min cell length (big-move rule, sph_equations3.cpp:330-335)
    float pseudo_factor;  // clamp(dt * pseudoViscosity, 0, 1)
    float thr_lo, thr_hi; // thr2 -+ band: below thr_lo certainly accepted, above thr_hi certainly rejected
    float xacc, xband;    // the same on x = 1 - d^2 / h^2 (list build): accepted when x > xacc, inside the band when x < xband
    int par;              // parity of the current grid epoch; integrate writes rebuild_flag[par ^ 1]
    int part;             // slab engines with the halo push: 0 = every block, 1 = the BOUNDARY blocks only (those holding slots of
                          // the first / last owned plane -- their results go to the neighbours and are wanted early), 2 = the rest
};

__device__ __forceinline__ int bbx_count(const StepParams &P){ return P.dyn ? P.dyn->n_own : P.n; }

// Boundary / interior split of a sweep over n slots in blocks of T (slots are sorted by cell, z slowest: the first owned
// plane is slots [0, n_first), the last one [n - n_last, n)).  bbx_part_blocks: how many blocks pass `part` has;
// bbx_part_block: the real block behind its v-th one.  A block that holds any boundary-plane slot is a boundary block, so
// an interior block never touches a ghost slot -- it needs nothing from the neighbours and can run while their halos travel.
struct PartMap { int tb0, tb1, nb; };
__device__ __forceinline__ PartMap bbx_part_map(const StepParams &P, int T, int n){
    PartMap m; m.nb = (n + T - 1) / T; m.tb0 = 0; m.tb1 = m.nb;
    if(P.part != 0){
        const int b0 = min(P.dyn->n_first, n), b1 = max(n - P.dyn->n_last, b0);
        m.tb0 = min((b0 + T - 1) / T, m.nb); m.tb1 = min(max(b1 / T, m.tb0), m.nb);
    }
    return m;
}
__device__ __forceinline__ int bbx_part_blocks(const StepParams &P, const PartMap &m){
    return P.part == 0 ? m.nb : (P.part == 1 ? m.tb0 + (m.nb - m.tb1) : m.tb1 - m.tb0);
}
__device__ __forceinline__ int bbx_part_block(const StepParams &P, const PartMap &m, int v){   // -1: past the end of the pass
    if(P.part == 0) return v < m.nb ? v : -1;
    if(P.part == 1){ const int b = v < m.tb0 ? v : m.tb1 + (v - m.tb0); return b < m.nb ? b : -1; }
    const int b = m.tb0 + v; return b < m.tb1 ? b : -1;
}

// ------------------------------------------------------------------------------------------- grid

// Grid::GetHashedPosition: FP64 arithmetic on the FP32-stored position, same expression order as the
// reference so that integer results are bit-exact for identical inputs. Returns -1 if outside.
__device__ __forceinline__ int bbx_hash(const DevGrid &g, float px, float py, float pz, int *ux, int *uy, int *uz){
    const float pf[3] = {px, py, pz};
    int u[3];
#pragma unroll
    for(int i = 0; i < 3; i++){
        double p = (double)pf[i];
        double eps = 0.0;
        if(fabs(p - g.min[i]) < 1e-8) eps = (double)0.0001f;
        else if(fabs(p - g.max[i]) < 1e-8) eps = -(double)0.0001f;
        p = __dadd_rn(p, eps);
        double dp = __ddiv_rn(__dsub_rn(p, g.min[i]), g.len[i]);
        u[i] = (int)floor(dp);
    }
    u[2] -= g.zoff; // local plane
    *ux = u[0]; *uy = u[1]; *uz = u[2];
    if(u[0] < 0 || u[0] >= g.n[0] || u[1] < 0 || u[1] >= g.n[1] || u[2] < 0 || u[2] >= g.n[2]) return -1;
    return u[0] + u[1] * g.n[0] + u[2] * g.plane;
}

// IsWithinStd(Distance(pi, pj), h) evaluated exactly as the reference does (FP64, sqrt then square,
// no FMA contraction): src/core/kernel.cpp:229-234, src/core/geometry.h:632-633.
__device__ __noinline__ bool bbx_within_std_exact(float4 a, float4 b, double h2){
    double x = __dsub_rn((double)a.x, (double)b.x);
    double y = __dsub_rn((double)a.y, (double)b.y);
    double z = __dsub_rn((double)a.z, (double)b.z);
    double s = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
    double d = __dsqrt_rn(s);
    double d2 = __dmul_rn(d, d);
    double of = __dsub_rn(d2, h2);
    return !(fabs(of) < 1e-8 || of > 0);
}

// Neighbour acceptance: FP32 fast path, FP64 only inside the guard band (|d2 - thr2| <= band).
__device__ __forceinline__ bool bbx_accept(const StepParams &P, float4 a, float4 b, float d2){
    float diff = d2 - P.thr2;
    if(fabsf(diff) <= P.band) return bbx_within_std_exact(a, b, P.h2_d);
    return diff < 0.f;
}

// ------------------------------------------------------------------------------------- colliders
// The collider response decides (penetrating / not, which face) on absolute epsilons of 1e-6 .. 1e-8;
// it runs in FP64 so that those decisions follow the FP64 reference (cost: O(100) flops per
// particle, twice per sub-step -- negligible next to the neighbour sums).

struct Vec3d { double x, y, z; };
__device__ __forceinline__ Vec3d v3(double x, double y, double z){ Vec3d r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ Vec3d operator+(Vec3d a, Vec3d b){ return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ Vec3d operator-(Vec3d a, Vec3d b){ return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ Vec3d operator*(double s, Vec3d a){ return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ double dot3(Vec3d a, Vec3d b){ return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double len3(Vec3d a){ return sqrt(dot3(a, a)); }
__device__ __forceinline__ double comp(Vec3d a, int i){ return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

__device__ __forceinline__ Vec3d xf_point(const double *m, Vec3d p){
    double xp = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    double yp = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    double zp = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    double wp = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    if(wp == 1) return v3(xp, yp, zp);
    double inv = 1.0 / wp;
    return v3(xp * inv, yp * inv, zp * inv);
}
// Transform::Normal uses the transposed inverse; for ObjectToWorld the inverse is WorldToObject.m
__device__ __forceinline__ Vec3d xf_normal(const double *minv, Vec3d n){
    return v3(minv[0] * n.x + minv[4] * n.y + minv[8] * n.z,
              minv[1] * n.x + minv[5] * n.y + minv[9] * n.z,
              minv[2] * n.x + minv[6] * n.y + minv[10] * n.z);
}

// Inside(vec3, Bounds3) with the reference's 1e-6 single-face slack (src/core/geometry.h:1773-1783)
__device__ __forceinline__ bool inside_bounds(Vec3d p, Vec3d lo, Vec3d hi){
    bool rv = (p.x >= lo.x && p.x <= hi.x && p.y >= lo.y && p.y <= hi.y && p.z >= lo.z && p.z <= hi.z);
    if(!rv){
        double ox = fmin(fabs(lo.x - p.x), fabs(hi.x - p.x));
        double oy = fmin(fabs(lo.y - p.y), fabs(hi.y - p.y));
        double oz = fmin(fabs(lo.z - p.z), fabs(hi.z - p.z));
        rv = (ox < 1e-6) || (oy < 1e-6) || (oz < 1e-6);
    }
    return rv;
}
__device__ __forceinline__ double clampd(double v, double lo, double hi){ return v < lo ? lo : (v > hi ? hi : v); }

// Box closest point in object space. Face order +x,+y,+z,-x,-y,-z with strict '<' (first wins).
__device__ __forceinline__ double box_query(const DevCollider &c, Vec3d pw, bool want_point, Vec3d *cpw, Vec3d *nw){
    Vec3d pl = xf_point(c.w2o, pw);
    Vec3d hi = v3(c.size[0] / 2.0, c.size[1] / 2.0, c.size[2] / 2.0);
    Vec3d lo = v3(-hi.x, -hi.y, -hi.z);
    Vec3d closest, normal = v3(1, 0, 0);
    bool neg = false;
    if(inside_bounds(pl, lo, hi)){
        // projections onto the six face planes; the distance to face k is |pl_k -+ h_k|
        double best = 0; int bi = 0;
#pragma unroll
        for(int i = 0; i < 6; i++){
            int ax = i % 3;
            double plane = i < 3 ? comp(hi, ax) : comp(lo, ax);
            double sgn = i < 3 ? 1.0 : -1.0;
            // Plane3::ClosestPoint: r - dot(r, n) n + point, r = p - point  (box.cpp:22-26)
            Vec3d pt = i < 3 ? hi : lo;
            Vec3d r = pl - pt;
            double d = sgn * comp(r, ax);
            Vec3d nn = v3(ax == 0 ? sgn : 0, ax == 1 ? sgn : 0, ax == 2 ? sgn : 0);
            Vec3d loc = (r - d * nn) + pt;
            Vec3d df = loc - pl;
            double l2 = dot3(df, df);
            if(i == 0 || l2 < best){ best = l2; bi = i; closest = loc; normal = nn; }
            (void)plane;
        }
        (void)bi;
        neg = true;
    }else{
        closest = v3(clampd(pl.x, lo.x, hi.x), clampd(pl.y, lo.y, hi.y), clampd(pl.z, lo.z, hi.z));
        Vec3d cov = pl - closest;
        double maxCos = cov.x; // dot(cov, (1,0,0))
#pragma unroll
        for(int i = 1; i < 6; i++){
            int ax = i % 3;
            double sgn = i < 3 ? 1.0 : -1.0;
            double cs = sgn * comp(cov, ax);
            if(cs > maxCos){ maxCos = cs; normal = v3(ax == 0 ? sgn : 0, ax == 1 ? sgn : 0, ax == 2 ? sgn : 0); }
        }
    }
    double dist = len3(closest - pl);
    if(neg) dist = -dist;
    if(want_point){
        if(c.reverse) normal = v3(-normal.x, -normal.y, -normal.z);
        *cpw = xf_point(c.o2w, closest);
        *nw = xf_normal(c.w2o, normal);
    }
    return dist;
}

__device__ __forceinline__ double lerpd(double a, double b, double t){ return (1 - t) * a + t * b; }

// FieldGrid::Sample (vertex centred, clamped barycentric indices)
__device__ __forceinline__ double sdf_sample(const DevCollider &c, Vec3d p){
    int ii[3], jj[3]; double w[3];
#pragma unroll
    for(int i = 0; i < 3; i++){
        double x = (comp(p, i) - c.sdf_origin[i]) / c.sdf_spacing[i];
        int high = c.sdf_res[i] - 1;
        double s = floor(x);
        int id = (int)s; double f;
        if(high == 0){ id = 0; f = 0; }
        else if(id < 0){ id = 0; f = 0; }
        else if(id > high - 1){ id = high - 1; f = 1; }
        else f = x - s;
        ii[i] = id; w[i] = f;
        jj[i] = min(id + 1, c.sdf_res[i] - 1);
    }
    const double *F = c.sdf_field; int rx = c.sdf_res[0], rxy = c.sdf_res[0] * c.sdf_res[1];
    double f000 = F[ii[0] + ii[1] * rx + ii[2] * rxy], f100 = F[jj[0] + ii[1] * rx + ii[2] * rxy];
    double f010 = F[ii[0] + jj[1] * rx + ii[2] * rxy], f110 = F[jj[0] + jj[1] * rx + ii[2] * rxy];
    double f001 = F[ii[0] + ii[1] * rx + jj[2] * rxy], f101 = F[jj[0] + ii[1] * rx + jj[2] * rxy];
    double f011 = F[ii[0] + jj[1] * rx + jj[2] * rxy], f111 = F[jj[0] + jj[1] * rx + jj[2] * rxy];
    double b0 = lerpd(lerpd(f000, f100, w[0]), lerpd(f010, f110, w[0]), w[1]);
    double b1 = lerpd(lerpd(f001, f101, w[0]), lerpd(f011, f111, w[0]), w[1]);
    return lerpd(b0, b1, w[2]);
}
// FieldGrid::Gradient: central differences with step = node spacing
__device__ __forceinline__ Vec3d sdf_gradient(const DevCollider &c, Vec3d p){
    double gx = (sdf_sample(c, v3(p.x + c.sdf_spacing[0], p.y, p.z)) - sdf_sample(c, v3(p.x - c.sdf_spacing[0], p.y, p.z))) * (1.0 / (2.0 * c.sdf_spacing[0]));
    double gy = (sdf_sample(c, v3(p.x, p.y + c.sdf_spacing[1], p.z)) - sdf_sample(c, v3(p.x, p.y - c.sdf_spacing[1], p.z))) * (1.0 / (2.0 * c.sdf_spacing[1]));
    double gz = (sdf_sample(c, v3(p.x, p.y, p.z + c.sdf_spacing[2])) - sdf_sample(c, v3(p.x, p.y, p.z - c.sdf_spacing[2]))) * (1.0 / (2.0 * c.sdf_spacing[2]));
    return v3(gx, gy, gz);
}

// DistanceTriangle (src/shapes/bvh.cpp:212-242): the edge projections are clamped to [0.0001, 1] as there
__device__ __forceinline__ Vec3d cross3(Vec3d a, Vec3d b){ return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ double edge_dist2(Vec3d e, Vec3d q){
    const double t = clampd(dot3(e, q) / dot3(e, e), 0.0001, 1.0);
    const Vec3d v = t * e - q;
    return dot3(v, v);
}
__device__ __forceinline__ double distance_triangle(const DevCollider &c, Vec3d p, int t){
    const int *ix = c.mesh_tris + 3 * (size_t)t;
    const double *pa_ = c.mesh_points + 3 * (size_t)ix[0], *pb_ = c.mesh_points + 3 * (size_t)ix[1], *pc_ = c.mesh_points + 3 * (size_t)ix[2];
    const Vec3d a = v3(pa_[0], pa_[1], pa_[2]), b = v3(pb_[0], pb_[1], pb_[2]), cc = v3(pc_[0], pc_[1], pc_[2]);
    const Vec3d ba = b - a, pa = p - a, cb = cc - b, pb = p - b, ac = a - cc, pc = p - cc;
    const Vec3d nor = cross3(ba, ac);
    auto sgn = [](double x){ return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); };
    const double s = sgn(dot3(cross3(ba, nor), pa)) + sgn(dot3(cross3(cb, nor), pb)) + sgn(dot3(cross3(ac, nor), pc));
    const double v = s < 2.0 ? fmin(fmin(edge_dist2(ba, pa), edge_dist2(cb, pb)), edge_dist2(ac, pc))
                             : dot3(nor, pa) * dot3(nor, pa) / dot3(nor, nor);
    return sqrt(v);
}
// Shape::MeshClosestDistance (BVHMeshClosestDistance, src/shapes/bvh.cpp:500-557): nearest triangle through the BVH; a
// subtree is entered only while its box is closer than the best distance so far, nearer child first
__device__ __noinline__ double mesh_closest_distance(const DevCollider &c, Vec3d p){
    double best = 3.1622776601683794E+18; // SqrtInfinity
    int stack[64]; int sp = 0; int node = 0;
    auto box_d2 = [&](const DevBvhNode &n){
        const double x = p.x - clampd(p.x, n.lo[0], n.hi[0]), y = p.y - clampd(p.y, n.lo[1], n.hi[1]), z = p.z - clampd(p.z, n.lo[2], n.hi[2]);
        return x * x + y * y + z * z;
    };
    while(node >= 0){
        const DevBvhNode &n = c.bvh[node];
        if(n.count > 0){
            for(int k = 0; k < n.count; k++){ const double d = distance_triangle(c, p, n.first + k); if(d < best) best = d; }
            node = sp > 0 ? stack[--sp] : -1;
        }else{
            const double b2 = best * best;
            const double dl = box_d2(c.bvh[n.left]), dr = box_d2(c.bvh[n.right]);
            const bool vl = dl < b2, vr = dr < b2;
            if(vl && vr){
                const int first = dl < dr ? n.left : n.right, second = dl < dr ? n.right : n.left;
                if(sp < 64) stack[sp++] = second;
                node = first;
            }else if(vl) node = n.left;
            else if(vr) node = n.right;
            else node = sp > 0 ? stack[--sp] : -1;
        }
    }
    return best;
}

__device__ __forceinline__ double closest_distance(const DevCollider &c, Vec3d p){
    if(c.type == BBX_COLLIDER_MESH) return mesh_closest_distance(c, p);
    if(c.type == BBX_COLLIDER_SPHERE){
        Vec3d pl = xf_point(c.w2o, p);
        return len3(pl) - c.radius;
    }else if(c.type == BBX_COLLIDER_BOX){
        return box_query(c, p, false, nullptr, nullptr);
    }
    return sdf_sample(c, p);
}

// ColliderSet3::ResolveCollision. Returns true when the particle was moved.
__device__ __noinline__ bool bbx_resolve_collision(const DevColliderSet &cs, double radius, double restitution,
                                                   float *px, float *py, float *pz, float *vx, float *vy, float *vz)
{
    Vec3d pos = v3(*px, *py, *pz);
    int target = -1; double minDistance = 3.4028234663852886e38; // Infinity == FLT_MAX
    for(int i = 0; i < cs.count; i++){
        // (a mesh only counts for points inside its bounds: Collider3::OptmizedClosestPointCheck, collider.cpp:113-121)
        if(cs.c[i].active && !(cs.c[i].type == BBX_COLLIDER_MESH && !inside_bounds(pos, v3(cs.c[i].mesh_lo[0], cs.c[i].mesh_lo[1], cs.c[i].mesh_lo[2]),
                                                                              v3(cs.c[i].mesh_hi[0], cs.c[i].mesh_hi[1], cs.c[i].mesh_hi[2])))){
            double d = fabs(closest_distance(cs.c[i], pos));
            if(d < minDistance){ target = i; minDistance = d; }
        }
    }
    if(target < 0) return false;
    const DevCollider &c = cs.c[target];
    Vec3d cp, nrm, cv = v3(0, 0, 0); double sd; bool inside;
    if(c.type == BBX_COLLIDER_SDF || c.type == BBX_COLLIDER_MESH){ // any shape with a grid: Shape::ClosestPointBySDF
        Vec3d tn = v3(0, 1, 0);
        Vec3d point = xf_point(c.w2o, pos);
        Vec3d tp = point;
        bool hasGradient = false;
        for(int it = 0; it < 5; it++){
            double sdf = sdf_sample(c, tp);
            if(fabs(sdf) < 0.001) break;
            tn = sdf_gradient(c, tp);
            double l = len3(tn);
            if(l > 0){ double inv = 1.0 / l; tn = v3(tn.x * inv, tn.y * inv, tn.z * inv); }
            tp = tp - sdf * tn;
            hasGradient = true;
        }
        if(!hasGradient){
            tn = sdf_gradient(c, tp);
            double l = len3(tn);
            if(l > 0){ double inv = 1.0 / l; tn = v3(tn.x * inv, tn.y * inv, tn.z * inv); }
        }
        sd = len3(point - tp);
        cp = xf_point(c.o2w, tp);
        nrm = xf_normal(c.w2o, tn);
        inside = sdf_sample(c, pos) < 0; // Shape::IsInside with a filled grid
    }else if(c.type == BBX_COLLIDER_SPHERE){
        Vec3d pl = xf_point(c.w2o, pos);
        sd = len3(pl) - c.radius;
        Vec3d N = v3(0, 1, 0);
        if(!(fabs(pl.x) < 1e-8 && fabs(pl.y) < 1e-8 && fabs(pl.z) < 1e-8)){
            double inv = 1.0 / len3(pl);
            N = v3(pl.x * inv, pl.y * inv, pl.z * inv);
        }
        Vec3d pN = c.radius * N;
        if(c.reverse) N = v3(-N.x, -N.y, -N.z);
        cp = xf_point(c.o2w, pN);
        nrm = xf_normal(c.w2o, N);
        // Shape::VelocityAt
        Vec3d q = cp - v3(c.o2w[3], c.o2w[7], c.o2w[11]);
        cv = v3(c.linvel[0] + (c.angvel[1] * q.z - c.angvel[2] * q.y),
                c.linvel[1] + (c.angvel[2] * q.x - c.angvel[0] * q.z),
                c.linvel[2] + (c.angvel[0] * q.y - c.angvel[1] * q.x));
        inside = (c.reverse != 0) == !(sd < 0);
    }else{
        sd = box_query(c, pos, true, &cp, &nrm);
        inside = (c.reverse != 0) == !(sd < 0);
    }
    // Collider3::IsPenetrating (collider.cpp:123-133)
    if(c.type == BBX_COLLIDER_MESH && !inside_bounds(pos, v3(c.mesh_lo[0], c.mesh_lo[1], c.mesh_lo[2]), v3(c.mesh_hi[0], c.mesh_hi[1], c.mesh_hi[2]))) return false;
    if(!(inside || fabs(sd) < radius)) return false;
    Vec3d tp = cp + radius * nrm;
    Vec3d vel = v3(*vx, *vy, *vz);
    Vec3d rel = vel - cv;
    double ndr = dot3(nrm, rel);
    Vec3d rn = ndr * nrm;
    Vec3d rt = rel - rn;
    if(ndr < 0){
        Vec3d dn = (-1.0 - restitution) * rn;
        rn = (-restitution) * rn;
        double rt2 = dot3(rt, rt);
        if(rt2 > 0){
            double scale = len3(dn) / sqrt(rt2);
            double fs = fmax(0.0, 1.0 - c.friction * scale);
            rt = fs * rt;
        }
        vel = rn + rt + cv;
        *vx = (float)vel.x; *vy = (float)vel.y; *vz = (float)vel.z;
    }
    *px = (float)tp.x; *py = (float)tp.y; *pz = (float)tp.z;
    return true;
}

// ------------------------------------------------------------------------- conservative FP32 pre-check
// FieldGrid::Sample on the FP32 copy of the field (same clamped-index convention as sdf_sample)
__device__ __forceinline__ float cull_sdf_sample(const DevCullCollider &c, float px, float py, float pz){
    const float pf[3] = {px, py, pz};
    int ii[3], jj[3]; float w[3];
#pragma unroll
    for(int i = 0; i < 3; i++){
        float x = (pf[i] - c.sdf_origin[i]) * c.sdf_inv_spacing[i];
        int high = c.sdf_res[i] - 1;
        float s = floorf(x);
        int id = (int)s; float f = x - s;
        if(high == 0 || id < 0){ id = 0; f = 0.f; }
        else if(id > high - 1){ id = high - 1; f = 1.f; }
        ii[i] = id; w[i] = f; jj[i] = min(id + 1, high);
    }
    const float *F = c.sdf_field32; int rx = c.sdf_res[0], rxy = c.sdf_res[0] * c.sdf_res[1];
    float f000 = __ldg(F + ii[0] + ii[1] * rx + ii[2] * rxy), f100 = __ldg(F + jj[0] + ii[1] * rx + ii[2] * rxy);
    float f010 = __ldg(F + ii[0] + jj[1] * rx + ii[2] * rxy), f110 = __ldg(F + jj[0] + jj[1] * rx + ii[2] * rxy);
    float f001 = __ldg(F + ii[0] + ii[1] * rx + jj[2] * rxy), f101 = __ldg(F + jj[0] + ii[1] * rx + jj[2] * rxy);
    float f011 = __ldg(F + ii[0] + jj[1] * rx + jj[2] * rxy), f111 = __ldg(F + jj[0] + jj[1] * rx + jj[2] * rxy);
    float a0 = f000 + w[0] * (f100 - f000), a1 = f010 + w[0] * (f110 - f010);
    float a2 = f001 + w[0] * (f101 - f001), a3 = f011 + w[0] * (f111 - f011);
    float b0 = a0 + w[1] * (a1 - a0), b1 = a2 + w[1] * (a3 - a2);
    return b0 + w[2] * (b1 - b0);
}

// true  = no active collider can be penetrated by a particle of this radius at p (certain, FP32 error
//         covered by cs.margin): ColliderSet3::ResolveCollision would leave position and velocity alone;
// false = undecided: run the exact FP64 response.
// Box / sphere: |signed distance| of the chosen collider decides (Collider3::IsPenetrating,
// collider.cpp:123-133); if every active collider is clear, so is whichever one is picked as the target.
// SDF grid: the projection moves the point by |f| per step and stops at |f| < 1e-3, so with L >= |grad f|
// the returned distance is >= (f(p) - 1e-3) / L (assumes the <= 5 Newton steps converge, which holds for
// any field close to a distance field; DESIGN.md states this).
__device__ __forceinline__ bool bbx_cull(const DevCullSet &cs, float px, float py, float pz, float radius){
    const float need = radius + cs.margin;
    for(int i = 0; i < cs.count; i++){
        const DevCullCollider &c = cs.c[i];
        if(!c.active) continue;
        float lx = px, ly = py, lz = pz;
        if(!c.identity){
            lx = fmaf(c.w2o[0], px, fmaf(c.w2o[1], py, fmaf(c.w2o[2], pz, c.w2o[3])));
            ly = fmaf(c.w2o[4], px, fmaf(c.w2o[5], py, fmaf(c.w2o[6], pz, c.w2o[7])));
            lz = fmaf(c.w2o[8], px, fmaf(c.w2o[9], py, fmaf(c.w2o[10], pz, c.w2o[11])));
        }
        float sd;
        if(c.type == BBX_COLLIDER_BOX){
            float qx = fabsf(lx) - c.half[0], qy = fabsf(ly) - c.half[1], qz = fabsf(lz) - c.half[2];
            // Inside(vec3, Bounds3) counts a point within 1e-6 of ANY face plane as inside (geometry.h:1773-1783,
            // mirrored by inside_bounds): near such a plane the box reports a tiny negative distance
            const float slab = 1.0e-6f + cs.margin;
            if(fminf(fabsf(qx), fminf(fabsf(qy), fabsf(qz))) < slab) return false;
            float ox = fmaxf(qx, 0.f), oy = fmaxf(qy, 0.f), oz = fmaxf(qz, 0.f);
            sd = sqrtf(fmaf(ox, ox, fmaf(oy, oy, oz * oz))) + fminf(fmaxf(qx, fmaxf(qy, qz)), 0.f);
        }else if(c.type == BBX_COLLIDER_SPHERE){
            sd = sqrtf(fmaf(lx, lx, fmaf(ly, ly, lz * lz))) - c.radius;
        }else{
            float s0 = cull_sdf_sample(c, lx, ly, lz);
            float sw = c.identity ? s0 : cull_sdf_sample(c, px, py, pz); // Shape::IsInside samples at the world position
            if(!(sw > cs.margin)) return false;
            if(!((s0 - 0.001f) > need * c.lipschitz)) return false;
            continue;
        }
        if(c.reverse){ if(!(sd < -need)) return false; }
        else{ if(!(sd > need)) return false; }
    }
    return true;
}
// certainly inside the domain (no clamp): sph_equations3.cpp:317-326
__device__ __forceinline__ bool bbx_inside_domain_certain(const DevCullSet &cs, float px, float py, float pz){
    return px > cs.dom_lo[0] && px < cs.dom_hi[0] && py > cs.dom_lo[1] && py < cs.dom_hi[1] && pz > cs.dom_lo[2] && pz < cs.dom_hi[2];
}

// ------------------------------------------------------------------------------ halo push (slab engines)
// Where the boundary-plane results of a sweep ALSO go: straight into the two neighbours' ghost slots
// (peer device memory: CUDA IPC mappings over NVLink, or plain pointers for slabs sharing a device).
// My first owned plane, slots [0, n_first), is the lower neighbour's upper ghost plane, its slots
// [lo_base, lo_base + n_first) with lo_base = that neighbour's owned count; my last owned plane, slots
// [hi_begin, n), is the upper neighbour's lower ghost plane, its slots [hi_begin - n, 0).
// Up to two arrays per kernel (x and v of the integration).  Inactive: n_first = 0, hi_begin = INT_MAX.
struct HaloDst {
    float4 *lo[2], *hi[2];
    int n_first, lo_base, hi_begin, n;
    int has_lo, has_hi; // a neighbour exists on that side and takes pushed results
};
// the counts live in DevState (no host round trip in the slab grid update): every kernel that pushes resolves them once
__device__ __forceinline__ HaloDst bbx_halo_resolve(HaloDst h, const DevState *st){
    if(h.has_lo | h.has_hi){
        h.n = st->n_own;
        h.n_first = h.has_lo ? st->n_first : 0; h.lo_base = st->peer_n[0];
        h.hi_begin = h.has_hi ? st->n_own - st->n_last : 0x7fffffff;
    }
    return h;
}
// float4 k of the `rec`-float4 record of slot i, array `which`
__device__ __forceinline__ void bbx_halo_store(const HaloDst &h, int which, int rec, int i, int k, float4 v){
    if(i < h.n_first) h.lo[which][(size_t)(h.lo_base + i) * rec + k] = v;
    if(i >= h.hi_begin) h.hi[which][(ptrdiff_t)(i - h.n) * rec + k] = v;
}
#define BBX_HALO_PHASES 6   // density, predict, pressure, integrate, grid: counts, grid: boundary planes
#define BBX_HALO_MAIL 4     // mailbox integers per side behind the flags: boundary-plane particles, owned particles, -, -
// After the kernels of a phase: tell both neighbours that my stores into their ghost slots are complete
// (stream order puts this after the producing kernels; the fence orders it after their stores system-wide).
__global__ void k_halo_signal(unsigned *lo_flag, unsigned *hi_flag, unsigned seq){
    if(threadIdx.x == 0){
        __threadfence_system();
        if(lo_flag) *(volatile unsigned *)lo_flag = seq;
        if(hi_flag) *(volatile unsigned *)hi_flag = seq;
    }
}
// Before the next phase reads ghost slots: wait until both neighbours have signalled this phase (bounded:
// ~30 s of polling -- a neighbour's host may be busy between two calls -- then the sticky device error
// BBX_ERR_COMM instead of a hang).
__global__ void k_halo_wait(const unsigned *from_lo, const unsigned *from_hi, unsigned seq, int *error){
    if(threadIdx.x == 0){
        const long long t0 = clock64();
        for(int side = 0; side < 2; side++){
            const volatile unsigned *f = side == 0 ? from_lo : from_hi;
            if(!f) continue;
            while((int)(*f - seq) < 0){
                if(clock64() - t0 > 60000000000ll){ *error = BBX_ERR_COMM; break; }
                __nanosleep(200);
            }
        }
        __threadfence_system();
    }
}

// warp helpers
__device__ __forceinline__ unsigned lanemask_lt(){ unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
