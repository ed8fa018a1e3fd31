// Mesh colliders on the host side of the facade (included by bubbles_api.h, namespace bbx):
//
//   reference                                             here
//   shape.h:275, shapes/bvh.cpp:51-56      MakeMesh       MakeMesh(points, triangles): ShapeMesh with its bounds
//   shapes/bvh.cpp:212-245                 DistanceTriangle (same operation order -> the same bits)
//   shapes/bvh.cpp:500-557                 BVHMeshClosestDistance -> MeshClosestDistance: the minimum over the triangles (the
//                                          reference's BVH only prunes; the minimum it finds is the minimum of the same values)
//   core/shape.cpp:358-377                 MeshIsPointInside: ray from the point towards the centre of the bounds, odd number
//                                          of crossings = inside
//   core/shape.cpp:411-431, 479-511        SetNodeSDFKernel + GenerateShapeSDF -> GenerateShapeSDF(shape, dx, margin): the
//                                          vertex-centred grid the collider set bakes for a mesh (|closest distance|, at least
//                                          1e-5, negative inside)
//
// The reference bakes the grid in a GPU kernel at set-up time; here it is a host loop at set-up time (nodes x triangles with
// a bounding-box cut-off -- meant for collider meshes of 1e3..1e5 triangles and grids of 1e4..1e6 nodes).  The engine gets
// the triangles (nearest-collider pick through its device BVH) and the grid (closest point, normal, inside test) through
// bbx_set_colliders as BBX_COLLIDER_MESH.  tests/test_mesh_sdf.py: the baked field equals, bit for bit, the one the
// unmodified reference bakes for the same mesh (tests/golden/mesh_collider.npz).
#pragma once

inline Float Sign(Float a){ int t = a < 0 ? -1 : 0; return a > 0 ? 1 : t; }   // geometry.h:125-128
inline vec3f Cross(const vec3f &a, const vec3f &b){                          // geometry.h:906-914
    return vec3f((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
inline Float Dot(const vec3f &a, const vec3f &b){ return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Float Dot2(const vec3f &a){ return Dot(a, a); }
inline Float Clamp(Float v, Float lo, Float hi){ if(v < lo) return lo; if(v > hi) return hi; return v; }

// DistanceTriangle (shapes/bvh.cpp:212-245)
inline Float DistanceTriangle(const vec3f &p, const vec3f &a, const vec3f &b, const vec3f &c){
    vec3f ba = b - a; vec3f pa = p - a;
    vec3f cb = c - b; vec3f pb = p - b;
    vec3f ac = a - c; vec3f pc = p - c;
    vec3f nor = Cross(ba, ac);
    auto mn = [](Float x, Float y){ return x < y ? x : y; };
    return std::sqrt(
        Sign(Dot(Cross(ba, nor), pa)) + Sign(Dot(Cross(cb, nor), pb)) + Sign(Dot(Cross(ac, nor), pc)) < 2.0 ?
            mn(mn(Dot2(ba * Clamp(Dot(ba, pa) / Dot2(ba), 0.0001, 1.0) - pa),
                  Dot2(cb * Clamp(Dot(cb, pb) / Dot2(cb), 0.0001, 1.0) - pb)),
               Dot2(ac * Clamp(Dot(ac, pc) / Dot2(ac), 0.0001, 1.0) - pc))
            : Dot(nor, pa) * Dot(nor, pa) / Dot2(nor));
}

// MakeMesh: world-space triangles (as Transform::Mesh leaves them in the reference), bounds = union of the triangle bounds
inline ShapePtr MakeMesh(const std::vector<vec3f> &points, const std::vector<int> &triangles, bool reverseOrientation = false){
    if(triangles.size() % 3 || triangles.empty()) throw std::invalid_argument("MakeMesh: triangles must hold 3 indices each");
    ShapePtr s = std::make_shared<Shape>(); s->type = ShapeMesh; s->reverseOrientation = reverseOrientation;
    s->meshPoints = points; s->meshIndices = triangles;
    const Float inf = std::numeric_limits<Float>::infinity();
    vec3f lo(inf, inf, inf), hi(-inf, -inf, -inf);
    for(int i : triangles){
        if(i < 0 || (size_t)i >= points.size()) throw std::invalid_argument("MakeMesh: vertex index out of range");
        const vec3f &q = points[(size_t)i];
        lo = vec3f(std::fmin(lo.x, q.x), std::fmin(lo.y, q.y), std::fmin(lo.z, q.z));
        hi = vec3f(std::fmax(hi.x, q.x), std::fmax(hi.y, q.y), std::fmax(hi.z, q.z));
    }
    s->meshBounds = Bounds3f(lo, hi);
    return s;
}

// closest distance to the surface; `hint` = a triangle to try first (the previous node's winner makes the cut-off bite at once)
inline Float MeshClosestDistance(const Shape &s, const vec3f &p, int *closest = nullptr, int hint = -1){
    const size_t nt = s.meshIndices.size() / 3;
    Float best = std::numeric_limits<Float>::infinity(); int arg = -1;
    auto tri = [&](size_t t){
        const vec3f &a = s.meshPoints[(size_t)s.meshIndices[3 * t]], &b = s.meshPoints[(size_t)s.meshIndices[3 * t + 1]], &c = s.meshPoints[(size_t)s.meshIndices[3 * t + 2]];
        // cut-off: a triangle whose bounding box is farther than the best distance so far cannot lower the minimum
        // (compared with a relative slack, so that rounding in this bound never hides a candidate)
        Float d2 = 0;
        for(int k = 0; k < 3; k++){
            const Float lo = std::fmin(a[k], std::fmin(b[k], c[k])), hi = std::fmax(a[k], std::fmax(b[k], c[k]));
            const Float e = p[k] < lo ? lo - p[k] : (p[k] > hi ? p[k] - hi : 0.0);
            d2 += e * e;
        }
        if(d2 * (1.0 - 1e-9) > best * best) return;
        const Float d = DistanceTriangle(p, a, b, c);
        if(d < best){ best = d; arg = (int)t; }
    };
    if(hint >= 0 && (size_t)hint < nt) tri((size_t)hint);
    for(size_t t = 0; t < nt; t++) if((int)t != hint) tri(t);
    if(closest) *closest = arg;
    return best;
}

// inside test by ray parity: the ray leaves the point towards the centre of the bounds (shape.cpp:358-377) and every crossing
// of the surface is counted.  Crossings are found triangle by triangle (Moeller-Trumbore in FP64); a ray that passes within
// 1e-9 (barycentric) of an edge or a vertex would count that crossing twice or not at all, so such a ray is abandoned and the
// count repeated along a slightly turned direction (a closed surface gives the same parity for every direction).
inline bool MeshIsPointInside(const Shape &s, const vec3f &p){
    const vec3f centre((s.meshBounds.pMin.x + s.meshBounds.pMax.x) * 0.5, (s.meshBounds.pMin.y + s.meshBounds.pMax.y) * 0.5, (s.meshBounds.pMin.z + s.meshBounds.pMax.z) * 0.5);
    vec3f d = centre - p;
    if(d.Length() < 1e-12) d = vec3f(1, 0, 0);
    const size_t nt = s.meshIndices.size() / 3;
    for(int attempt = 0; attempt < 16; attempt++){
        vec3f dir = d * (1.0 / d.Length());
        int hits = 0; bool grazing = false;
        for(size_t t = 0; t < nt && !grazing; t++){
            const vec3f &a = s.meshPoints[(size_t)s.meshIndices[3 * t]], &b = s.meshPoints[(size_t)s.meshIndices[3 * t + 1]], &c = s.meshPoints[(size_t)s.meshIndices[3 * t + 2]];
            const vec3f e1 = b - a, e2 = c - a, pv = Cross(dir, e2);
            const Float det = Dot(e1, pv);
            if(std::fabs(det) < 1e-300) continue;                       // parallel to the triangle's plane
            const Float inv = 1.0 / det;
            const vec3f tv = p - a;
            const Float u = Dot(tv, pv) * inv;
            const vec3f qv = Cross(tv, e1);
            const Float v = Dot(dir, qv) * inv, w = 1.0 - u - v, tt = Dot(e2, qv) * inv;
            const Float eps = 1e-9;
            if(u < -eps || v < -eps || w < -eps || tt < -eps) continue;  // clearly beside or behind
            if(u < eps || v < eps || w < eps || tt < eps){ grazing = true; break; }
            hits++;
        }
        if(!grazing) return (hits % 2) != 0;
        // turn the direction a little (deterministic) and count again
        d = vec3f(d.x + 1.0e-3 * (attempt + 1) * d.Length(), d.y - 0.7e-3 * (attempt + 1) * d.Length(), d.z + 0.4e-3 * (attempt + 1) * d.Length());
    }
    return false;
}

// GenerateShapeSDF(Shape *, dx, margin), shape.cpp:479-511 + SetNodeSDFKernel :411-431
inline void GenerateShapeSDF(Shape *shape, Float dx = 0.01, Float margin = 0.1){
    if(shape->type != ShapeMesh) throw std::invalid_argument("GenerateShapeSDF: not a mesh shape");
    vec3f lo = shape->meshBounds.pMin, hi = shape->meshBounds.pMax;
    vec3f sc(std::fabs(hi.x - lo.x), std::fabs(hi.y - lo.y), std::fabs(hi.z - lo.z));
    lo = lo - sc * margin; hi = hi + sc * margin;
    const Float width = std::fabs(hi.x - lo.x), height = std::fabs(hi.y - lo.y), depth = std::fabs(hi.z - lo.z);
    const int res = (int)std::ceil(width / dx);
    dx = width / (Float)res;
    const int ry = (int)std::ceil(res * height / width), rz = (int)std::ceil(res * depth / width);
    // vertex centred: Get1DLengthFor(resolution) = resolution + 1 nodes per axis (grid.h:1211-1238)
    shape->sdfResolution[0] = res + 1; shape->sdfResolution[1] = ry + 1; shape->sdfResolution[2] = rz + 1;
    shape->sdfSpacing = dx; shape->sdfOrigin = lo;
    shape->sdfBounds = Bounds3f(lo, vec3f(lo.x + dx * (Float)res, lo.y + dx * (Float)ry, lo.z + dx * (Float)rz));
    shape->sdfField.resize((size_t)(res + 1) * (ry + 1) * (rz + 1));
    size_t k = 0; int hint = -1;
    for(int z = 0; z <= rz; z++) for(int y = 0; y <= ry; y++) for(int x = 0; x <= res; x++){
        const vec3f p(lo.x + dx * x, lo.y + dx * y, lo.z + dx * z);           // FieldGrid::GetDataPosition, grid.h:1010-1026
        const Float d = MeshClosestDistance(*shape, p, &hint, hint);
        const bool interior = MeshIsPointInside(*shape, p);
        const Float psd = std::fabs(d) < 0.00001 ? 0.00001 : std::fabs(d);   // Max(Absf(d), 0.00001)
        shape->sdfField[k++] = interior ? -psd : psd;
    }
}
