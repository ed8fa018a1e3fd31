// bubbles_api.h -- C++ host facade over the bbx C ABI (include/bbx.h) with the class / function names of
// felpzOliveira/Bubbles, so that a scene script ports by switching includes:
//
//   reference (src/...)                                   here (namespace bbx)
//   core/pcisph_solver.h:56-83   PciSphSolver3             PciSphSolver3   {Initialize, Setup, SetColliders, SetViscosityCoefficient,
//                                                                           Advance, GetSphSolverData, GetSphParticleSet, GetAdvanceTime,
//                                                                           GetParticleCount, ComputeDelta}
//   core/sph_solver.h:76-96      SphSolver3                SphSolver3
//   solvers/sph_solver3.cpp:124  DefaultSphSolverData3     DefaultSphSolverData3
//   core/particle.h:611-720      ParticleSetBuilder3, SphParticleSet3FromBuilder
//   core/emitter.h:39-128        VolumeParticleEmitter3, VolumeParticleEmitterSet3
//   core/shape.h:273-286         MakeBox, MakeSphere, MakeSDFShape; MakeMesh + GenerateShapeSDF (mesh_sdf.h here)
//   core/collider.h:113-122      ColliderSetBuilder3
//   core/util.cpp:269            UtilBuildGridForDomain
//   third/serializer.h:36        SerializerSaveSphDataSet3 (text frames bbtool reads), SERIALIZER_* flags
//   core/pcisph_solver.h:95      PciSphRunSimulation3 (the run loop, without the viewer)
//   core/transform_sequence.h    TransformSequence, QuaternionSequence, InterpolatedTransform, Quaternion (transform_sequence.h here)
//   core/shape.cpp:290-298       Shape::Update / SetVelocities + PciSphSolver3::UpdateCollider (hands the change to the engine)
//
// Differences by design: objects are ordinary C++ values / unique_ptrs (the reference bump-allocates from a
// managed-memory arena and never frees, src/cuda/memory.h); errors throw bbx::Error instead of
// getchar() + exit(0) (src/cuda/cutil.cpp:15-26); all device work happens inside libbbx.so -- this header is
// host-only and needs no nvcc.  Particle data lives on the device between steps; Advance() refreshes the
// host copies of positions / velocities / densities (what the reference's callbacks read through managed
// memory).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/bbx.h"

namespace bbx {

typedef double Float; // src/core/geometry.h:68-69
const Float WaterDensity = 1000.0;
const Float Pi = 3.14159265358979323846;

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error("bbx error " + std::to_string(c) + ": " + m), code(c) {}
};
inline void Check(int rc){ if(rc != BBX_OK) throw Error(rc, bbx_last_error()); }

struct vec3f {
    Float x, y, z;
    vec3f() : x(0), y(0), z(0) {}
    explicit vec3f(Float a) : x(a), y(a), z(a) {}
    vec3f(Float a, Float b, Float c) : x(a), y(b), z(c) {}
    Float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    vec3f operator+(const vec3f &o) const { return vec3f(x + o.x, y + o.y, z + o.z); }
    vec3f operator-(const vec3f &o) const { return vec3f(x - o.x, y - o.y, z - o.z); }
    vec3f operator*(Float s) const { return vec3f(x * s, y * s, z * s); }
    Float Length() const { return std::sqrt(x * x + y * y + z * z); }
};
inline vec3f operator*(Float s, const vec3f &v){ return v * s; }

struct Bounds3f {
    vec3f pMin, pMax;
    Bounds3f() {}
    Bounds3f(const vec3f &a, const vec3f &b)
        : pMin(std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)), pMax(std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)) {}
    Float ExtentOn(int i) const { return std::fabs(pMax[i] - pMin[i]); }
};

// 4x4 transform with its inverse, row-major (src/core/transform.h:399-468)
struct Transform {
    Float m[4][4], mInv[4][4];
    Transform(){ for(int i = 0; i < 4; i++) for(int j = 0; j < 4; j++) m[i][j] = mInv[i][j] = (i == j) ? 1.0 : 0.0; }
    vec3f Point(const vec3f &p) const {
        Float xp = m[0][0] * p.x + m[0][1] * p.y + m[0][2] * p.z + m[0][3];
        Float yp = m[1][0] * p.x + m[1][1] * p.y + m[1][2] * p.z + m[1][3];
        Float zp = m[2][0] * p.x + m[2][1] * p.y + m[2][2] * p.z + m[2][3];
        Float wp = m[3][0] * p.x + m[3][1] * p.y + m[3][2] * p.z + m[3][3];
        if(wp == 1) return vec3f(xp, yp, zp);
        return vec3f(xp, yp, zp) * (1.0 / wp);
    }
    vec3f InversePoint(const vec3f &p) const { Transform t; std::memcpy(t.m, mInv, sizeof(m)); std::memcpy(t.mInv, m, sizeof(m)); return t.Point(p); }
};
inline Transform Translate(Float x, Float y, Float z){ // analytic inverse, src/core/transform.cpp:286-292
    Transform t; t.m[0][3] = x; t.m[1][3] = y; t.m[2][3] = z; t.mInv[0][3] = -x; t.mInv[1][3] = -y; t.mInv[2][3] = -z; return t;
}
inline Transform Translate(const vec3f &v){ return Translate(v.x, v.y, v.z); }
inline Transform Scale(Float x, Float y, Float z){
    Transform t; t.m[0][0] = x; t.m[1][1] = y; t.m[2][2] = z; t.mInv[0][0] = 1 / x; t.mInv[1][1] = 1 / y; t.mInv[2][2] = 1 / z; return t;
}
// Transform products, Rotate, Quaternion, InterpolatedTransform, TransformSequence, QuaternionSequence (keyframed collider
// motion; src/core/transform_sequence.h, quaternion.h)
#include "transform_sequence.h"

// ------------------------------------------------------------------------------------------ shapes
enum ShapeType { ShapeSphere = BBX_COLLIDER_SPHERE, ShapeBox = BBX_COLLIDER_BOX, ShapeSDF = BBX_COLLIDER_SDF, ShapeMesh = BBX_COLLIDER_MESH };

struct Shape {
    ShapeType type = ShapeBox;
    Transform ObjectToWorld;
    bool reverseOrientation = false;
    Float radius = 0, sizex = 0, sizey = 0, sizez = 0;
    vec3f linearVelocity, angularVelocity;
    // Shape::Update / SetVelocities (src/core/shape.cpp:290-298): a scene script moves the collider between frames (e.g. with a
    // TransformSequence); PciSphSolver3::UpdateCollider hands the new state to the engine
    void Update(const Transform &toWorld){ ObjectToWorld = toWorld; }
    void SetVelocities(const vec3f &vel, const vec3f &angular){ linearVelocity = vel; angularVelocity = angular; }
    // baked SDF (FieldGrid3f, vertex centred): node counts, spacing, position of node (0, 0, 0), x-fastest values
    int sdfResolution[3] = {0, 0, 0};
    Float sdfSpacing = 0;
    vec3f sdfOrigin;
    std::vector<Float> sdfField;
    Bounds3f sdfBounds;
    // triangle mesh (ShapeMesh; world space): vertices, 3 vertex indices per triangle, bounds (mesh_sdf.h)
    std::vector<vec3f> meshPoints;
    std::vector<int> meshIndices;
    Bounds3f meshBounds;

    Bounds3f GetBounds() const {
        if(type == ShapeSDF) return sdfBounds;
        if(type == ShapeMesh) return meshBounds;
        vec3f h = type == ShapeSphere ? vec3f(radius) : vec3f(sizex / 2, sizey / 2, sizez / 2);
        // transformed corners (Transform::operator()(Bounds3f))
        Bounds3f b(ObjectToWorld.Point(vec3f(-h.x, -h.y, -h.z)), ObjectToWorld.Point(vec3f(-h.x, -h.y, -h.z)));
        for(int k = 1; k < 8; k++){
            vec3f c = ObjectToWorld.Point(vec3f((k & 1) ? h.x : -h.x, (k & 2) ? h.y : -h.y, (k & 4) ? h.z : -h.z));
            b = Bounds3f(vec3f(std::fmin(b.pMin.x, c.x), std::fmin(b.pMin.y, c.y), std::fmin(b.pMin.z, c.z)),
                         vec3f(std::fmax(b.pMax.x, c.x), std::fmax(b.pMax.y, c.y), std::fmax(b.pMax.z, c.z)));
        }
        return b;
    }
    // Shape::SignedDistance (src/core/shape.cpp:275-288) = -|d| inside, |d| outside, d = ClosestDistance.  Box:
    // BoxClosestDistance (src/shapes/box.cpp:74-109) with the reference's Inside(point, bounds)
    // (src/core/geometry.h:1773-1783): a point outside the box still counts as inside when it lies within 1e-6 of ANY
    // face plane -- the emitter keeps such lattice points, so the rule is mirrored here.
    Float SignedDistance(const vec3f &p) const {
        vec3f q = ObjectToWorld.InversePoint(p);
        Float d;
        if(type == ShapeSphere) d = q.Length() - radius;
        else if(type == ShapeBox){
            const vec3f hi(sizex / 2.0, sizey / 2.0, sizez / 2.0), lo(-hi.x, -hi.y, -hi.z);
            bool in = q.x >= lo.x && q.x <= hi.x && q.y >= lo.y && q.y <= hi.y && q.z >= lo.z && q.z <= hi.z;
            if(!in){
                const Float ox = std::fmin(std::fabs(lo.x - q.x), std::fabs(hi.x - q.x));
                const Float oy = std::fmin(std::fabs(lo.y - q.y), std::fabs(hi.y - q.y));
                const Float oz = std::fmin(std::fabs(lo.z - q.z), std::fabs(hi.z - q.z));
                in = ox < 1e-6 || oy < 1e-6 || oz < 1e-6;
            }
            if(in){ // distance to the nearest of the six face planes, negative
                Float best = std::fabs(q.x - hi.x);
                const Float c[5] = {std::fabs(q.y - hi.y), std::fabs(q.z - hi.z), std::fabs(q.x - lo.x), std::fabs(q.y - lo.y), std::fabs(q.z - lo.z)};
                for(Float v : c) if(v < best) best = v;
                d = -best;
            }else{
                const vec3f cl(std::fmin(std::fmax(q.x, lo.x), hi.x), std::fmin(std::fmax(q.y, lo.y), hi.y), std::fmin(std::fmax(q.z, lo.z), hi.z));
                d = (q - cl).Length();
            }
        }else d = SampleSDF(p);
        return reverseOrientation ? -d : d;
    }
    // FieldGrid::Sample (src/core/grid.h:1093-1131): trilinear, clamped indices
    Float SampleSDF(const vec3f &p) const {
        int ii[3], jj[3]; Float w[3];
        for(int a = 0; a < 3; a++){
            Float x = (p[a] - sdfOrigin[a]) / sdfSpacing; int high = sdfResolution[a] - 1;
            Float s = std::floor(x); int id = (int)s; Float f;
            if(high == 0 || id < 0){ id = 0; f = 0; }else if(id > high - 1){ id = high - 1; f = 1; }else f = x - s;
            ii[a] = id; w[a] = f; jj[a] = id + 1 < sdfResolution[a] ? id + 1 : sdfResolution[a] - 1;
        }
        const int rx = sdfResolution[0], rxy = sdfResolution[0] * sdfResolution[1];
        auto F = [&](int x, int y, int z){ return sdfField[(size_t)x + (size_t)y * rx + (size_t)z * rxy]; };
        auto L = [](Float a, Float b, Float t){ return (1 - t) * a + t * b; };
        Float b0 = L(L(F(ii[0], ii[1], ii[2]), F(jj[0], ii[1], ii[2]), w[0]), L(F(ii[0], jj[1], ii[2]), F(jj[0], jj[1], ii[2]), w[0]), w[1]);
        Float b1 = L(L(F(ii[0], ii[1], jj[2]), F(jj[0], ii[1], jj[2]), w[0]), L(F(ii[0], jj[1], jj[2]), F(jj[0], jj[1], jj[2]), w[0]), w[1]);
        return L(b0, b1, w[2]);
    }
};
typedef std::shared_ptr<Shape> ShapePtr;

inline ShapePtr MakeBox(const Transform &toWorld, const vec3f &size, bool reverseOrientation = false){
    ShapePtr s = std::make_shared<Shape>(); s->type = ShapeBox; s->ObjectToWorld = toWorld;
    s->sizex = size.x; s->sizey = size.y; s->sizez = size.z; s->reverseOrientation = reverseOrientation; return s;
}
inline ShapePtr MakeSphere(const Transform &toWorld, Float radius, bool reverseOrientation = false){
    ShapePtr s = std::make_shared<Shape>(); s->type = ShapeSphere; s->ObjectToWorld = toWorld; s->radius = radius;
    s->reverseOrientation = reverseOrientation; return s;
}
// MakeSDFShape(bounds, sdf) (src/core/shape.h:281-286): bakes an analytic signed-distance function on the
// vertex grid of Shape::InitSDFShape (shape.h:201-230: bounds grown by `margin`, node spacing ~dx)
inline ShapePtr MakeSDFShape(const Bounds3f &bounds, const std::function<Float(vec3f)> &sdf, Float dx = 0.01, Float margin = 0.1){
    ShapePtr s = std::make_shared<Shape>(); s->type = ShapeSDF;
    vec3f lo = bounds.pMin, hi = bounds.pMax;
    vec3f sc(std::fabs(hi.x - lo.x), std::fabs(hi.y - lo.y), std::fabs(hi.z - lo.z));
    lo = lo - sc * margin; hi = hi + sc * margin;
    Float width = std::fabs(hi.x - lo.x), height = std::fabs(hi.y - lo.y), depth = std::fabs(hi.z - lo.z);
    int res = (int)std::ceil(width / dx); dx = width / (Float)res;
    int ry = (int)std::ceil(res * height / width), rz = (int)std::ceil(res * depth / width);
    s->sdfResolution[0] = res + 1; s->sdfResolution[1] = ry + 1; s->sdfResolution[2] = rz + 1;
    s->sdfSpacing = dx; s->sdfOrigin = lo; s->sdfBounds = Bounds3f(lo, hi);
    s->sdfField.resize((size_t)(res + 1) * (ry + 1) * (rz + 1));
    size_t k = 0;
    for(int z = 0; z <= rz; z++) for(int y = 0; y <= ry; y++) for(int x = 0; x <= res; x++)
        s->sdfField[k++] = sdf(vec3f(lo.x + dx * x, lo.y + dx * y, lo.z + dx * z));
    return s;
}

// mesh colliders: MakeMesh, DistanceTriangle, MeshClosestDistance, MeshIsPointInside, GenerateShapeSDF
#include "mesh_sdf.h"

// --------------------------------------------------------------------------------------- colliders
struct ColliderSet3 {
    std::vector<ShapePtr> shapes;
    std::vector<Float> friction;
    std::vector<bool> active;
    int nColiders() const { return (int)shapes.size(); }
    void SetActive(int which, bool on){ active.at(which) = on; }
    // ColliderSet3::GenerateSDFs -> Collider3::GenerateSDFs -> GenerateShapeSDF (src/core/collider.cpp:146-149, 238-264): bake the grid of every mesh collider that has none yet
    void GenerateSDFs(Float dx = 0.01, Float margin = 0.1){ for(auto &s : shapes) if(s->type == ShapeMesh && s->sdfField.empty()) GenerateShapeSDF(s.get(), dx, margin); }
    std::vector<bbx_collider> ToABI() const {
        std::vector<bbx_collider> out(shapes.size());
        for(size_t i = 0; i < shapes.size(); i++){
            const Shape &s = *shapes[i]; bbx_collider &b = out[i];
            std::memset(&b, 0, sizeof(b));
            b.type = (int)s.type; b.reverse_orientation = s.reverseOrientation ? 1 : 0; b.active = active[i] ? 1 : 0; b.friction = friction[i];
            for(int r = 0; r < 4; r++) for(int c = 0; c < 4; c++){ b.object_to_world[4 * r + c] = s.ObjectToWorld.m[r][c]; b.world_to_object[4 * r + c] = s.ObjectToWorld.mInv[r][c]; }
            b.size[0] = s.sizex; b.size[1] = s.sizey; b.size[2] = s.sizez; b.radius = s.radius;
            for(int k = 0; k < 3; k++){ b.linear_velocity[k] = s.linearVelocity[k]; b.angular_velocity[k] = s.angularVelocity[k]; }
            if(s.type == ShapeMesh){   // the triangles pick the nearest collider, the baked grid answers the rest (GenerateSDFs first)
                if(s.sdfField.empty()) throw std::invalid_argument("mesh collider without its SDF grid: call ColliderSet3::GenerateSDFs / GenerateShapeSDF first");
                static_assert(sizeof(vec3f) == 3 * sizeof(double), "vec3f must be three packed doubles");
                b.mesh_vertices = (int)s.meshPoints.size(); b.mesh_triangles = (int)(s.meshIndices.size() / 3);
                b.mesh_points = reinterpret_cast<const double *>(s.meshPoints.data()); b.mesh_indices = s.meshIndices.data();
            }
            if(s.type == ShapeSDF || s.type == ShapeMesh){
                for(int k = 0; k < 3; k++){ b.sdf_resolution[k] = s.sdfResolution[k]; b.sdf_spacing[k] = s.sdfSpacing; b.sdf_origin[k] = s.sdfOrigin[k]; }
                b.sdf_field = s.sdfField.data();
            }
        }
        return out;
    }
};
struct ColliderSetBuilder3 {
    std::shared_ptr<ColliderSet3> set = std::make_shared<ColliderSet3>();
    void AddCollider3(const ShapePtr &shape, Float frictionCoefficient = 0.0){ // MakeCollider3 defaults, src/core/collider.cpp:340-348
        set->shapes.push_back(shape); set->friction.push_back(frictionCoefficient); set->active.push_back(true);
    }
    std::shared_ptr<ColliderSet3> GetColliderSet(){ return set; }
};

// --------------------------------------------------------------------------------------------- grid
struct Grid3 { bbx_grid_desc desc; Bounds3f GetBounds() const { return Bounds3f(vec3f(desc.min[0], desc.min[1], desc.min[2]), vec3f(desc.max[0], desc.max[1], desc.max[2])); }
               int GetCellCount() const { return desc.total; } };
inline std::shared_ptr<Grid3> UtilBuildGridForDomain(const Bounds3f &domain, Float spacing, Float spacingScale){
    auto g = std::make_shared<Grid3>();
    double lo[3] = {domain.pMin.x, domain.pMin.y, domain.pMin.z}, hi[3] = {domain.pMax.x, domain.pMax.y, domain.pMax.z};
    Check(bbx_grid_for_domain(lo, hi, spacing, spacingScale, &g->desc));
    return g;
}

// ---------------------------------------------------------------------------------------- particles
struct ParticleSetBuilder3 {
    std::vector<vec3f> positions, velocities;
    int AddParticle(const vec3f &pos, const vec3f &vel = vec3f(0)){ positions.push_back(pos); velocities.push_back(vel); return 1; }
    void SetVelocityForAll(const vec3f &vel){ for(auto &v : velocities) v = vel; }
    int GetParticleCount() const { return (int)positions.size(); }
    void Commit(){}
};

// ParticleSet3 + SphParticleSet3 (src/core/particle.h:154-213, 480-608): host mirror of the device state
struct ParticleSet3 {
    std::vector<vec3f> positions, velocities;
    std::vector<Float> densities;
    std::vector<Float> v0s;        // boundary layer of each particle (ParticleSet3::v0s, particle.h:170): 0 = interior
    std::vector<vec3f> normals;    // filled by a boundary / normal pass of the caller (src/boundaries/*, out of scope here); zeros otherwise
    Float mass = 0;
    int GetParticleCount() const { return (int)positions.size(); }
    vec3f GetParticlePosition(int i) const { return positions[i]; }
    vec3f GetParticleVelocity(int i) const { return velocities[i]; }
    Float GetParticleDensity(int i) const { return densities[i]; }
    vec3f GetParticleNormal(int i) const { return (size_t)i < normals.size() ? normals[(size_t)i] : vec3f(); }
    void SetParticleNormal(int i, const vec3f &n){ if(normals.size() < positions.size()) normals.resize(positions.size()); normals[(size_t)i] = n; }
    Float GetParticleV0(int i) const { return (size_t)i < v0s.size() ? v0s[(size_t)i] : 0.0; }
    void SetParticleV0(int i, Float v){ if(v0s.size() < positions.size()) v0s.resize(positions.size(), 0.0); v0s[(size_t)i] = v; }
    Float GetMass() const { return mass; }
};
struct SphParticleSet3 {
    ParticleSet3 set;
    Float targetSpacing = 0.1, kernelRadiusOverSpacing = 2.0, targetDensity = WaterDensity;
    int reservedSize = 0;          // ContinuousParticleSetBuilder3: room for particles emitted during the run
    bbx_engine *engine = nullptr;  // set by the solver's Setup: appended particles go to the device as well
    ParticleSet3 *GetParticleSet(){ return &set; }
    void SetRelativeKernelRadius(Float r){ kernelRadiusOverSpacing = r; }
    void SetTargetSpacing(Float s){ targetSpacing = s; }
    Float GetTargetSpacing() const { return targetSpacing; }
    Float GetKernelRadius() const { return kernelRadiusOverSpacing * targetSpacing; }
};
inline std::shared_ptr<SphParticleSet3> SphParticleSet3FromBuilder(ParticleSetBuilder3 *b){
    auto s = std::make_shared<SphParticleSet3>();
    s->set.positions = b->positions; s->set.velocities = b->velocities; s->set.densities.assign(b->positions.size(), 0.0);
    return s;
}

// The re-emission rule of ContinuousParticleSetBuilder3::MapGridEmit (src/core/grid.h:1367-1407) as a pure function:
// for every mapped cell (ascending id) whose CURRENT chain has room (< 100 = MaximumParticlesPerBucket), its first
// min(100 - size, template size) template positions are re-emitted unless a particle of the current chain lies
// within d.  cell_count / cell_order = the chains (bbx_export_cells), positions = current positions by id.
inline std::vector<vec3f> MapGridEmitCandidates(const std::map<unsigned, std::vector<vec3f>> &mapped, int total_cells,
        const int *cell_count, const int *cell_order, const vec3f *positions, Float d){
    std::vector<int> start((size_t)total_cells + 1, 0);
    for(int c = 0; c < total_cells; c++) start[(size_t)c + 1] = start[(size_t)c] + cell_count[c];
    std::vector<vec3f> out;
    for(const auto &kv : mapped){
        const unsigned h = kv.first; const std::vector<vec3f> &tmpl = kv.second;
        if((int)h >= total_cells) continue;
        const int size = cell_count[h];
        if(size >= BBX_MAX_NEIGHBORS) continue;
        const int toInsert = std::min(BBX_MAX_NEIGHBORS - size, (int)tmpl.size());
        for(int i = 0; i < toInsert; i++){
            bool can_add = true;
            for(int j = 0; j < size && can_add; j++) if((positions[cell_order[start[h] + j]] - tmpl[(size_t)i]).Length() < d) can_add = false;
            if(can_add) out.push_back(tmpl[(size_t)i]);
        }
    }
    return out;
}

// ContinuousParticleSetBuilder3 (src/core/grid.h:1246-1450): a particle set with reserved room; AddParticle queues,
// Commit appends (ids continue from the current count) -- on the device the new particles join the tail of their
// cells' chains exactly like DistributeByParticleList (bbx_append_particles).  MapGrid (after the solver's Setup)
// remembers the chains' positions as the emission template; MapGridEmit re-emits them between frames where there is
// room (the chains come from the engine: bbx_export_cells; the positions are the solver's host mirror, so call it
// after Advance).
struct ContinuousParticleSetBuilder3 {
    std::vector<vec3f> positions, velocities;
    std::shared_ptr<SphParticleSet3> particleSet = std::make_shared<SphParticleSet3>();
    int maxNumOfParticles;
    int reemitions = 0; bool reemitOnce = false;
    std::map<unsigned, std::vector<vec3f>> mappedPositions;
    int mappedCells = 0;
    explicit ContinuousParticleSetBuilder3(int maxParticles = 2500000, bool _reemitOnce = false)
        : maxNumOfParticles(maxParticles > 1 ? maxParticles : 1), reemitOnce(_reemitOnce) { particleSet->reservedSize = maxNumOfParticles; }
    void SetKernelRadius(Float){}
    int AddParticle(const vec3f &pos, const vec3f &vel = vec3f(0)){
        if(particleSet->set.GetParticleCount() + (int)positions.size() + 1 > maxNumOfParticles) return 0;
        positions.push_back(pos); velocities.push_back(vel); return 1;
    }
    void Commit(){
        if(positions.empty()) return;
        ParticleSet3 &ps = particleSet->set;
        if(particleSet->engine) Check(bbx_append_particles(particleSet->engine, (int)positions.size(), positions.data(), velocities.data(), BBX_F64));
        ps.positions.insert(ps.positions.end(), positions.begin(), positions.end());
        ps.velocities.insert(ps.velocities.end(), velocities.begin(), velocities.end());
        ps.densities.resize(ps.positions.size(), 0.0);
        positions.clear(); velocities.clear();
    }
    int GetParticleCount() const { return particleSet->set.GetParticleCount(); }
    // chains of the engine: counts per cell + concatenated ids
    void Chains(const Grid3 &grid, std::vector<int> *count, std::vector<int> *order) const {
        if(!particleSet->engine) throw Error(BBX_ERR_INVALID, "MapGrid / MapGridEmit need the solver's Setup() first");
        count->assign((size_t)grid.desc.total, 0); order->assign((size_t)std::max(1, GetParticleCount()), 0);
        Check(bbx_export_cells(particleSet->engine, count->data(), order->data()));
    }
    template<typename G> void MapGrid(const G &grid){ // G: Grid3 or a smart pointer to it
        const Grid3 &gr = DerefGrid(grid);
        std::vector<int> count, order; Chains(gr, &count, &order);
        mappedPositions.clear(); mappedCells = gr.desc.total;
        size_t at = 0;
        for(int c = 0; c < gr.desc.total; c++){
            if(count[(size_t)c] > 0){
                std::vector<vec3f> &v = mappedPositions[(unsigned)c];
                for(int j = 0; j < count[(size_t)c]; j++) v.push_back(particleSet->set.positions[(size_t)order[at + (size_t)j]]);
            }
            at += (size_t)count[(size_t)c];
        }
        mapped = &gr;
    }
    int MapGridEmit(const std::function<vec3f(const vec3f &)> &velocity, Float d = 0.02){
        if((reemitions > 0 && reemitOnce) || !mapped) return 0;
        // the per-cell test runs on the device (bbx_query_cells): template points go up, 8 bytes per point come back -- the
        // chains are not downloaded (the reference walks them on the host, grid.h:1367-1407)
        std::vector<int> cells; std::vector<vec3f> pts;
        for(const auto &kv : mappedPositions){
            if((int)kv.first >= mapped->desc.total) continue;
            for(const vec3f &p : kv.second){ cells.push_back((int)kv.first); pts.push_back(p); }
        }
        std::vector<int> size(cells.size()), blocked(cells.size());
        if(!particleSet->engine) throw Error(BBX_ERR_INVALID, "MapGridEmit needs the solver's Setup() first");
        static_assert(sizeof(vec3f) == 3 * sizeof(double), "vec3f must be 3 packed doubles");
        Check(bbx_query_cells(particleSet->engine, (int)cells.size(), cells.data(), (const double *)pts.data(), d, size.data(), blocked.data()));
        std::vector<vec3f> add;
        for(size_t k = 0; k < cells.size();){
            size_t e = k; while(e < cells.size() && cells[e] == cells[k]) e++;       // the template of one cell
            const int sz = size[k];
            if(sz >= 0 && sz < BBX_MAX_NEIGHBORS){
                const size_t toInsert = std::min<size_t>((size_t)(BBX_MAX_NEIGHBORS - sz), e - k);
                for(size_t i = 0; i < toInsert; i++) if(!blocked[k + i]) add.push_back(pts[k + i]);
            }
            k = e;
        }
        int added = 0;
        for(const vec3f &p : add){ if(!AddParticle(p, velocity(p))) break; added++; }
        if(added > 0) Commit();
        reemitions++;
        return added;
    }
  private:
    const Grid3 *mapped = nullptr;
    static const Grid3 &DerefGrid(const Grid3 &g){ return g; }
    static const Grid3 &DerefGrid(const Grid3 *g){ return *g; }
    static const Grid3 &DerefGrid(const std::shared_ptr<Grid3> &g){ return *g; }
};
inline std::shared_ptr<SphParticleSet3> SphParticleSet3FromContinuousBuilder(ContinuousParticleSetBuilder3 *b){ b->Commit(); return b->particleSet; }

// ----------------------------------------------------------------------------------------- emitters
// BccLatticePointGenerator::ForEach (src/generator/bcclattice.cpp:5-36)
template<typename F> inline void BccLatticeForEach(const Bounds3f &b, Float spacing, F &&fn){
    const Float half = spacing / 2;
    const Float ex = b.ExtentOn(0), ey = b.ExtentOn(1), ez = b.ExtentOn(2);
    bool shifted = false;
    for(int k = 0; k * half <= ez; k++){
        const Float off = shifted ? half : 0.0, z = k * half + b.pMin.z;
        for(int j = 0; j * spacing + off <= ey; j++){
            const Float y = j * spacing + off + b.pMin.y;
            for(int i = 0; i * spacing + off <= ex; i++) if(!fn(vec3f(i * spacing + off + b.pMin.x, y, z))) return;
        }
        shifted = !shifted;
    }
}
// VolumeParticleEmitter3 (src/core/emitter.h:39-69, emitter.cpp:242-330): BCC lattice over `bound`, each point
// jittered by 0.5 * jitter * spacing * SampleSphere(rand, rand) (libc rand(), like the reference), kept when
// shape->SignedDistance(target) <= 0.
struct VolumeParticleEmitter3 {
    ShapePtr shape; Bounds3f bound; Float spacing; vec3f initVel; Float jitter = 0; int maxParticles = 0x7fffffff; int emittedParticles = 0;
    std::function<bool(const vec3f &)> validator;
    VolumeParticleEmitter3(const ShapePtr &s, const Bounds3f &b, Float sp, const vec3f &v = vec3f(0)) : shape(s), bound(b), spacing(sp), initVel(v) {}
    void SetJitter(Float j){ jitter = j < 0 ? 0 : (j > 1 ? 1 : j); }
    void SetValidator(std::function<bool(const vec3f &)> f){ validator = std::move(f); }
    template<typename Builder> void Emit(Builder *builder){ // ParticleSetBuilder3 or ContinuousParticleSetBuilder3
        const Float maxJitter = 0.5 * jitter * spacing;
        BccLatticeForEach(bound, spacing, [&](const vec3f &point) -> bool {
            if(validator && !validator(point)) return true;
            // vec2f u(rand_float(), rand_float()) (emitter.cpp:272): g++ evaluates the arguments right to left, so the FIRST
            // draw is u[1]; kept that way so that an emission reproduces the reference's particles
            const float u1 = rand() / (RAND_MAX + 1.f), u0 = rand() / (RAND_MAX + 1.f);
            const Float usqrt = 2 * std::sqrt((Float)u1 * (1 - (Float)u1)), utheta = 2 * Pi * (Float)u0;
            const vec3f target = point + maxJitter * vec3f(std::cos(utheta) * usqrt, std::sin(utheta) * usqrt, 1 - 2 * (Float)u1);
            if(shape->SignedDistance(target) <= 0){
                if(emittedParticles >= maxParticles) return false;
                if(!builder->AddParticle(target, initVel)) return false; // builder full: the emission stops (emitter.cpp:283-289)
                emittedParticles++;
            }
            return true;
        });
        builder->Commit(); // the reference's _Emit commits per emitter (emitter.cpp:300-303)
    }
};
struct VolumeParticleEmitterSet3 {
    std::vector<VolumeParticleEmitter3 *> emitters;
    void AddEmitter(VolumeParticleEmitter3 *e){ emitters.push_back(e); }
    void SetJitter(Float j){ for(auto *e : emitters) e->SetJitter(j); }
    template<typename Builder> void Emit(Builder *b){ for(auto *e : emitters) e->Emit(b); b->Commit(); }
};

// ------------------------------------------------------------------------------------------ solvers
// SphSolverData3 constants (src/core/sph_solver.h:31-50); the arrays live in the engine
struct SphSolverData3 {
    bbx_config cfg;
    std::shared_ptr<SphParticleSet3> sphpSet;
    std::shared_ptr<Grid3> domain;
    std::shared_ptr<ColliderSet3> collider;
};
inline std::shared_ptr<SphSolverData3> DefaultSphSolverData3(bool with_gravity = true){
    auto d = std::make_shared<SphSolverData3>();
    Check(bbx_config_default(&d->cfg, with_gravity ? 1 : 0));
    return d;
}

class SolverBase3 {
  protected:
    std::shared_ptr<SphSolverData3> data;
    bbx_engine *engine = nullptr;
    int solverKind;
    Float stepInterval = 0;
    explicit SolverBase3(int kind) : solverKind(kind) {}
    void Pull(){
        ParticleSet3 &ps = data->sphpSet->set;
        if(ps.positions.empty()) return;
        static_assert(sizeof(vec3f) == 3 * sizeof(double), "vec3f must be 3 packed doubles");
        Check(bbx_download(engine, BBX_POSITION, ps.positions.data(), BBX_F64));
        Check(bbx_download(engine, BBX_VELOCITY, ps.velocities.data(), BBX_F64));
        Check(bbx_download(engine, BBX_DENSITY, ps.densities.data(), BBX_F64));
    }
  public:
    SolverBase3(const SolverBase3 &) = delete;
    SolverBase3 &operator=(const SolverBase3 &) = delete;
    ~SolverBase3(){ if(engine){ if(data && data->sphpSet) data->sphpSet->engine = nullptr; bbx_destroy(engine); } }
    void Initialize(const std::shared_ptr<SphSolverData3> &d){ data = d; }
    void Setup(Float targetDensity, Float targetSpacing, Float relativeRadius, const std::shared_ptr<Grid3> &domain,
               const std::shared_ptr<SphParticleSet3> &pSet, int maxParticles = 0){
        if(!data) throw Error(BBX_ERR_INVALID, "Initialize() before Setup()");
        data->sphpSet = pSet; data->domain = domain;
        pSet->targetDensity = targetDensity; pSet->targetSpacing = targetSpacing; pSet->kernelRadiusOverSpacing = relativeRadius;
        bbx_config &c = data->cfg;
        c.spacing = targetSpacing; c.kernel_scale = relativeRadius; c.target_density = targetDensity; c.grid = domain->desc;
        const int n = pSet->set.GetParticleCount();
        c.max_particles = maxParticles > n ? maxParticles : (n > 0 ? n : 1);
        if(pSet->reservedSize > c.max_particles) c.max_particles = pSet->reservedSize;
        if(engine){ bbx_destroy(engine); engine = nullptr; }
        Check(bbx_create(&c, &engine));
        Check(bbx_get_mass(engine, &pSet->set.mass));
        if(n > 0) Check(bbx_set_particles(engine, n, pSet->set.positions.data(), pSet->set.velocities.data(), BBX_F64));
        pSet->engine = engine;
        if(data->collider) SetColliders(data->collider);
    }
    void SetColliders(const std::shared_ptr<ColliderSet3> &colliders){
        data->collider = colliders;
        if(!engine) return; // applied by Setup
        std::vector<bbx_collider> abi = colliders->ToABI();
        Check(bbx_set_colliders(engine, (int)abi.size(), abi.data()));
    }
    std::shared_ptr<ColliderSet3> GetColliders(){ return data->collider; }
    // after Shape::Update / SetVelocities / ColliderSet3::SetActive on collider `which` (the reference's kernels read the
    // managed-memory Shape live; here the change is handed over explicitly): bbx_update_collider + bbx_set_collider_active
    void UpdateCollider(int which){
        if(!engine || !data->collider) return;
        std::vector<bbx_collider> abi = data->collider->ToABI();
        Check(bbx_update_collider(engine, which, &abi.at((size_t)which)));
        Check(bbx_set_collider_active(engine, which, abi[(size_t)which].active));
    }
    // the reference reads SphSolverData3 live every sub-step: setters called after Setup() reach the engine too
    void SetParam(int param, double value){ if(engine) Check(bbx_set_param(engine, param, value)); }
    void SetViscosityCoefficient(Float v){ data->cfg.viscosity = v > 0 ? v : 0; SetParam(BBX_PARAM_VISCOSITY, data->cfg.viscosity); }
    SphSolverData3 *GetSphSolverData(){ return data.get(); }
    SphParticleSet3 *GetSphParticleSet(){ return data->sphpSet.get(); }
    Float GetKernelRadius(){ return data->sphpSet->GetKernelRadius(); }
    Float GetAdvanceTime() const { return stepInterval; }
    int GetParticleCount(){ int n = 0; Check(bbx_particle_count(engine, &n)); return n; }
    bbx_engine *Engine(){ return engine; }
    // Advance(timeIntervalInSeconds): CFL sub-stepping on the device, then the host copies are refreshed
    int Advance(Float timeIntervalInSeconds){
        int substeps = 0; float ms = 0;
        Check(bbx_advance(engine, timeIntervalInSeconds, solverKind, &substeps, &ms));
        stepInterval = ms;
        Pull();
        return substeps;
    }
    // fixed-dt sub-steps (AdvanceTimeStep), without the CFL loop
    void AdvanceTimeStep(Float dt, int count = 1){ Check(bbx_step_many(engine, dt, solverKind, count)); Pull(); }
    bbx_step_stats Stats(){ bbx_step_stats s; Check(bbx_stats(engine, &s)); return s; }
};
class PciSphSolver3 : public SolverBase3 {
  public:
    PciSphSolver3() : SolverBase3(BBX_SOLVER_PCISPH) {}
    Float ComputeDelta(Float timeIntervalInSeconds){ double d = 0; Check(bbx_get_delta(engine, timeIntervalInSeconds, &d)); return d; }
    // 1: the reference's effective behaviour (one predict-correct iteration); 0: iterate to tolerance
    void SetReferenceCompat(bool on){ data->cfg.pcisph_reference_compat = on ? 1 : 0; SetParam(BBX_PARAM_REFERENCE_COMPAT, on ? 1 : 0); }
};
class SphSolver3 : public SolverBase3 {
  public:
    SphSolver3() : SolverBase3(BBX_SOLVER_SPH) {}
    void SetPseudoViscosityCoefficient(Float v){ data->cfg.pseudo_viscosity = v; SetParam(BBX_PARAM_PSEUDO_VISCOSITY, v); }
};

// --------------------------------------------------------------------------------------- serializer
enum { SERIALIZER_POSITION = 0x01, SERIALIZER_VELOCITY = 0x02, SERIALIZER_DENSITY = 0x04, SERIALIZER_BOUNDARY = 0x08,
       SERIALIZER_NORMAL = 0x10, SERIALIZER_MASS = 0x20, SERIALIZER_LAYERS = 0x40, SERIALIZER_XYZ = 0x80,
       SERIALIZER_RULE_BOUNDARY_EXCLUSIVE = 0x100 };   // src/third/serializer.h:7-17
inline std::string SerializerStringFromFlags(int flags){   // serializer.cpp:48-59
    std::string s;
    if(flags & SERIALIZER_POSITION) s += "p";
    if(flags & SERIALIZER_VELOCITY) s += "v";
    if(flags & SERIALIZER_DENSITY) s += "d";
    if(flags & SERIALIZER_MASS) s += "m";
    if(flags & SERIALIZER_BOUNDARY) s += "b";
    if(flags & SERIALIZER_NORMAL) s += "n";
    if(flags & SERIALIZER_LAYERS) s += "l";
    if(flags & SERIALIZER_RULE_BOUNDARY_EXCLUSIVE) s += "o";
    return s;
}
// SaveSphParticleSet + PushParticleSetToFile (src/third/serializer.cpp:812-921): the "FluidBegin ... DataEnd / FluidEnd" text
// frame, one line per particle in id order, columns p v d m b n, "%g" / "%d"; appended to `filename` like the reference
// ("a+").  `boundary` = the per-particle boundary layer a classification pass produced (src/boundaries/*: the caller's, out of
// scope here); SERIALIZER_RULE_BOUNDARY_EXCLUSIVE keeps only the particles with boundary > 0; normals come from
// ParticleSet3::normals.  The quirks of the reference are kept: without a boundary vector the b column is dropped with a
// warning while the format string still says "b", and the exclusive rule without a vector writes nothing.
inline void SerializerSaveSphDataSet3(SphSolverData3 *data, const char *filename, int flags, std::vector<int> *boundary = nullptr){
    ParticleSet3 *ps = data->sphpSet->GetParticleSet();
    FILE *fp = std::fopen(filename, "a+");
    if(!fp){ std::printf("Error: Failed to open %s\n", filename); return; }
    int pCount = ps->GetParticleCount();
    if((flags & SERIALIZER_RULE_BOUNDARY_EXCLUSIVE) && boundary){
        pCount = 0;
        for(size_t i = 0; i < boundary->size(); i++) pCount += boundary->at(i) > 0 ? 1 : 0;
    }else if(flags & SERIALIZER_RULE_BOUNDARY_EXCLUSIVE){
        std::printf("Invalid configuration for Serialized particles\n");
        std::fclose(fp);
        return;
    }
    std::fprintf(fp, "FluidBegin\n\t\"Type\" particles\n\t\"Count\" %d\n\t\"Format\" %s\n\t\"Spacing\" %g\n\tDataBegin\n",
                 pCount, SerializerStringFromFlags(flags).c_str(), data->sphpSet->GetTargetSpacing());
    int logged = 0;
    for(int i = 0; i < ps->GetParticleCount(); i++){
        int sp = 0;
        const int boundary_value = boundary ? boundary->at((size_t)i) : 0;
        if((flags & SERIALIZER_RULE_BOUNDARY_EXCLUSIVE) && !boundary_value) continue;
        std::fprintf(fp, "\t\t");
        if(flags & SERIALIZER_POSITION){ vec3f p = ps->GetParticlePosition(i); std::fprintf(fp, "%g %g %g", p.x, p.y, p.z); sp = 1; }
        if(flags & SERIALIZER_VELOCITY){ vec3f v = ps->GetParticleVelocity(i); std::fprintf(fp, sp ? " %g %g %g" : "%g %g %g", v.x, v.y, v.z); sp = 1; }
        if(flags & SERIALIZER_DENSITY){ std::fprintf(fp, sp ? " %g" : "%g", ps->GetParticleDensity(i)); sp = 1; }
        if(flags & SERIALIZER_MASS){ std::fprintf(fp, sp ? " %g" : "%g", ps->GetMass()); sp = 1; }
        if(flags & SERIALIZER_BOUNDARY){
            if(!boundary && !logged){ std::printf("Warning: Not a valid boundary given\n"); logged = 1; }
            else if(boundary){ std::fprintf(fp, sp ? " %d" : "%d", boundary_value); sp = 1; }
        }
        if(flags & SERIALIZER_NORMAL){ vec3f n = ps->GetParticleNormal(i); std::fprintf(fp, sp ? " %g %g %g" : "%g %g %g", n.x, n.y, n.z); sp = 1; }
        std::fprintf(fp, "\n");
    }
    std::fprintf(fp, "\tDataEnd\nFluidEnd\n");
    std::fclose(fp);
}

// ---- frame reader: SerializerLoadParticles3 / SerializerLoadSphDataSet3 / SerializerLoadPoints3
// (src/third/serializer.cpp:84-261, 444-577).  Numbers are parsed the way the reference parses them (ParseFloat,
// src/third/obj_loader.cpp:270-278 = tinyobjloader's tryParseDouble: decimal digits accumulated in a double, the
// fraction digit by digit times 10^-k, a power-of-ten exponent applied as ldexp(m * 5^e, e)), NOT with strtod: the
// values loaded here are bit-identical to what bbtool and the reference's loaders see.
inline bool ParseDecimal(const char *s, const char *end, double *out){
    static const double frac_lut[8] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
    if(s >= end) return false;
    double m = 0.0; int e10 = 0; bool neg = false;
    const char *c = s;
    if(*c == '+' || *c == '-'){ neg = (*c == '-'); c++; }
    else if(!(*c >= '0' && *c <= '9')) return false;
    int digits = 0;
    while(c != end && *c >= '0' && *c <= '9'){ m = m * 10 + (int)(*c - '0'); c++; digits++; }
    if(digits == 0) return false;
    bool has_exp = false;
    if(c != end){
        if(*c == '.'){
            c++;
            int k = 1;
            while(c != end && *c >= '0' && *c <= '9'){ m += (int)(*c - '0') * (k < 8 ? frac_lut[k] : std::pow(10.0, -k)); k++; c++; }
            has_exp = (c != end) && (*c == 'e' || *c == 'E');
        }else has_exp = (*c == 'e' || *c == 'E');
    }
    if(has_exp){
        c++;
        bool eneg = false;
        if(c != end && (*c == '+' || *c == '-')){ eneg = (*c == '-'); c++; }
        else if(!(c != end && *c >= '0' && *c <= '9')) return false;   // an empty exponent is not a number
        int ed = 0;
        while(c != end && *c >= '0' && *c <= '9'){ e10 = e10 * 10 + (int)(*c - '0'); c++; ed++; }
        if(ed == 0) return false;
        if(eneg) e10 = -e10;
    }
    *out = (neg ? -1 : 1) * (e10 ? std::ldexp(m * std::pow(5.0, e10), e10) : m);
    return true;
}
inline Float ParseFloat(const char **token){
    *token += std::strspn(*token, " \t");
    const char *end = *token + std::strcspn(*token, " \t\r");
    double v = 0; ParseDecimal(*token, end, &v);
    *token = end;
    return v;
}
inline vec3f ParseV3(const char **token){ Float a = ParseFloat(token), b = ParseFloat(token), c = ParseFloat(token); return vec3f(a, b, c); }
inline int SerializerFlagsFromString(const char *spec){
    int flags = 0;
    for(const char *p = spec; *p; p++){
        switch(*p | 0x20){
            case 'p': flags |= SERIALIZER_POSITION; break; case 'v': flags |= SERIALIZER_VELOCITY; break;
            case 'd': flags |= SERIALIZER_DENSITY; break;  case 'm': flags |= SERIALIZER_MASS; break;
            case 'b': flags |= SERIALIZER_BOUNDARY; break; case 'n': flags |= SERIALIZER_NORMAL; break;
            case 'l': flags |= SERIALIZER_LAYERS; break;   case 'o': flags |= SERIALIZER_RULE_BOUNDARY_EXCLUSIVE; break;
            case 'z': flags |= SERIALIZER_XYZ; break;
            default: std::printf("Unknown flag argument %c\n", *p); return -1;
        }
    }
    return flags;
}
struct SerializedParticle { vec3f position, velocity, normal; Float density = 0, mass = 0; int boundary = 0; };
// header of the first fluid section: "Count", "Format"; leaves the stream behind its "DataBegin" line
inline int SerializerFindFluidSection(std::istream &is, std::string &format){
    std::string line; int count = 0; bool in_region = false;
    while(std::getline(is, line)){
        if(!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ls(line); std::string tok;
        while(ls >> tok){
            if(!in_region){ if(tok == "FluidBegin") in_region = true; continue; }
            if(tok == "\"Count\""){ std::string v; if(ls >> v){ const char *t = v.c_str(); count = (int)ParseFloat(&t); } }
            else if(tok == "\"Format\""){ ls >> format; }
            else if(tok == "DataBegin") return count;
            else if(tok == "FluidEnd") in_region = false;
        }
    }
    return count;
}
inline int SerializerLoadParticles3(std::vector<SerializedParticle> *pSet, const char *filename, int &flags){
    std::ifstream ifs(filename);
    if(!ifs){ std::printf("Could not open file %s\n", filename); return -1; }
    std::string format;
    const int expected = SerializerFindFluidSection(ifs, format);
    if(!format.empty()) flags = SerializerFlagsFromString(format.c_str());
    pSet->clear(); pSet->reserve(expected > 0 ? expected : 0);
    std::string line; bool found_end = false;
    while(std::getline(ifs, line)){
        if(!line.empty() && line.back() == '\r') line.pop_back();
        const char *token = line.c_str();
        token += std::strspn(token, " \t");
        if(token[0] == '\0' || token[0] == '#') continue;
        if(std::strstr(token, "DataEnd")){ found_end = true; break; }
        SerializedParticle q;
        auto skip = [&](){ while(*token == ' ' || *token == '\t' || *token == '\r') token++; };
        if(flags & SERIALIZER_POSITION){ q.position = ParseV3(&token); skip(); }
        if(flags & SERIALIZER_VELOCITY){ q.velocity = ParseV3(&token); skip(); }
        if(flags & SERIALIZER_DENSITY){ q.density = ParseFloat(&token); skip(); }
        if(flags & SERIALIZER_MASS){ q.mass = ParseFloat(&token); skip(); }
        if(flags & SERIALIZER_BOUNDARY){
            q.boundary = (int)ParseFloat(&token); skip();
            if((flags & SERIALIZER_RULE_BOUNDARY_EXCLUSIVE) && !q.boundary) continue;
        }
        if(flags & SERIALIZER_NORMAL){ q.normal = ParseV3(&token); skip(); }
        pSet->push_back(q);
    }
    if(!found_end) throw Error(BBX_ERR_INVALID, "Unterminated fluid description, missing 'DataEnd'");
    return (int)pSet->size();
}
inline int SerializerLoadSphDataSet3(ParticleSetBuilder3 *builder, const char *filename, int &flags, std::vector<int> *boundary = nullptr){
    std::vector<SerializedParticle> ps;
    const int n = SerializerLoadParticles3(&ps, filename, flags);
    for(int i = 0; i < n; i++){ builder->AddParticle(ps[i].position, ps[i].velocity); if(boundary) boundary->push_back(ps[i].boundary); }
    return n;
}
inline void SerializerLoadPoints3(std::vector<vec3f> *points, const char *filename, int &flags){
    std::vector<SerializedParticle> ps;
    const int n = SerializerLoadParticles3(&ps, filename, flags);
    points->clear();
    for(int i = 0; i < n; i++) points->push_back(ps[i].position);
}

// Shape::BoxSerialize / SphereSerialize (src/shapes/box.cpp:37-57, sphere.cpp:11-31): the "ShapeBegin ... ShapeEnd"
// block bbtool reads (stream formatting of the reference: 6 significant digits); other shape kinds have no text form.
inline std::string ShapeSerialize(const Shape &s){
    if(s.type != ShapeBox && s.type != ShapeSphere) return std::string();
    std::stringstream ss;
    ss << "ShapeBegin\n";
    if(s.type == ShapeBox){
        ss << "\t\"Type\" box" << std::endl;
        ss << "\t\"Length\" " << s.sizex << " " << s.sizey << " " << s.sizez << std::endl;
    }else{
        ss << "\t\"Type\" sphere" << std::endl;
        ss << "\t\"Radius\" " << s.radius << std::endl;
    }
    ss << "\t\"Transform\" ";
    for(int i = 0; i < 4; i++) for(int j = 0; j < 4; j++){
        if(i == 3 && j == 3) ss << s.ObjectToWorld.m[i][j]; else ss << s.ObjectToWorld.m[i][j] << " ";
    }
    ss << std::endl;
    ss << "ShapeEnd";
    return ss.str();
}
// UtilGetBoundaryState (src/core/util.h:708-725): the boundary column = the positive v0 values
inline int UtilGetBoundaryState(ParticleSet3 *pSet, std::vector<int> *boundaries){
    int n = 0; boundaries->clear();
    for(int i = 0; i < pSet->GetParticleCount(); i++){
        int b = 0; int v0 = (int)pSet->GetParticleV0(i);
        if(v0 > 0){ b = v0; n++; }
        boundaries->push_back(b);
    }
    return n;
}
// UtilSaveSimulation3 (src/core/util.h:296-328): the file is rewritten with the shape blocks of the active obstacle
// colliders -- every collider but the LAST, which the reference takes to be the domain -- followed by the particle
// block.  (A boundary vector, when the caller has one, goes through SerializerSaveSphDataSet3's last argument.)
inline void UtilSaveSimulation3(ColliderSet3 *colliders, SphSolverData3 *data, const char *filename, int flags){
    std::remove(filename);
    FILE *fp = std::fopen(filename, "a+");
    if(!fp){ std::printf("Failed to open file %s\n", filename); return; }
    std::stringstream ss;
    if(colliders) for(int i = 0; i < colliders->nColiders() - 1; i++) if(colliders->active[i]) ss << ShapeSerialize(*colliders->shapes[i]) << std::endl;
    std::fprintf(fp, "%s", ss.str().c_str());
    std::fclose(fp);
    std::vector<int> boundaries;
    UtilGetBoundaryState(data->sphpSet->GetParticleSet(), &boundaries);
    SerializerSaveSphDataSet3(data, filename, flags, &boundaries);
}
template<typename Solver, typename ParticleAccessor>
inline void UtilSaveSimulation3(Solver *solver, ParticleAccessor *, const char *filename, int flags){
    UtilSaveSimulation3(solver->GetColliders().get(), solver->GetSphSolverData(), filename, flags);
}

// PciSphRunSimulation3 / UtilRunSimulation3 (src/core/util.h:539-599) without the viewer: callback(step) before the
// first frame with step = 0, then after every Advance; it returns 0 to stop.
template<typename Solver>
inline void RunSimulation3(Solver *solver, Float targetInterval, const std::function<int(int)> &callback){
    int step = 0;
    if(!callback(step)) return;
    for(;;){ solver->Advance(targetInterval); step++; if(!callback(step)) break; }
}
inline void PciSphRunSimulation3(PciSphSolver3 *solver, Float targetInterval, const std::function<int(int)> &callback){ RunSimulation3(solver, targetInterval, callback); }

// Wavefront .obj files for mesh colliders: LoadObj, MakeMesh(path | mesh, toWorld)
#include "obj_loader.h"

} // namespace bbx
