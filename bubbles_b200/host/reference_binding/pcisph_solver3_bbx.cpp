// Reference-side binding: the file a Bubbles maintainer drops into the reference tree as
// src/solvers/pcisph_solver3_bbx.cpp and compiles INSTEAD OF the stepping functions of src/solvers/pcisph_solver3.cpp
// (-DBUBBLES_USE_BBX=ON).  It keeps PciSphSolver3's members and re-routes Setup / SetColliders / SetViscosityCoefficient /
// Advance through the bbx C ABI (include/bbx.h); everything above the solver (scene scripts, emitters, builders, the run
// loop, the serializer, bbtool) keeps working on the reference's own objects.
//
// Not part of libbbx.  The repository's reference build recipe COMPILES it against the unmodified reference headers whenever
// the reference sources are present (compile check only; tests/test_abi.py looks for the object), so the binding cannot
// drift from either side's declarations.
#include <pcisph_solver.h>   // reference: src/core/pcisph_solver.h:56-83
#include <collider.h>        // reference: src/core/collider.h:55-100
#include <shape.h>           // reference: src/core/shape.h:140-180
#include <bbx.h>
#include <vector>
#include <cstring>
#include <cstdio>

static bbx_engine *g_bbx = nullptr;   // one engine per solver (a real patch stores it in PciSphSolver3)

static void BbxRowMajor(const Transform &t, double out[16], double inv[16]){
    for(int i = 0; i < 4; i++) for(int j = 0; j < 4; j++){ out[4 * i + j] = t.m.m[i][j]; inv[4 * i + j] = t.mInv.m[i][j]; }
}

static void BbxReport(const char *what){ printf("bbx: %s failed: %s\n", what, bbx_last_error()); }

// PciSphSolver3::Setup (src/solvers/pcisph_solver3.cpp:103-146): allocations, mass, delta denominator, first distribution
void PciSphSolver3::Setup(Float targetDensity, Float targetSpacing, Float relativeRadius,
                          Grid3 *domain, SphParticleSet3 *pSet)
{
    SphSolverData3 *sph = solverData->sphData;
    sph->sphpSet = pSet; sph->domain = domain;              // the reference's own bookkeeping stays
    bbx_config cfg; bbx_config_default(&cfg, /*with_gravity=*/1);
    cfg.spacing = targetSpacing; cfg.kernel_scale = relativeRadius; cfg.target_density = targetDensity;
    cfg.viscosity = sph->viscosity; cfg.drag = sph->dragCoefficient; cfg.pseudo_viscosity = sph->pseudoViscosity;
    cfg.eos_exponent = sph->eosExponent; cfg.sound_speed = sph->soundSpeed;
    cfg.negative_pressure_scale = sph->negativePressureScale;
    cfg.pcisph_max_iterations = (int)maxIterations; cfg.pcisph_max_density_error_ratio = maxErrorDensity;
    ParticleSet3 *ps = pSet->GetParticleSet();
    cfg.max_particles = ps->GetReservedSize();
    vec3f p0 = domain->GetBounds().pMin, p1 = domain->GetBounds().pMax;
    int res[3] = {(int)domain->usizes[0], (int)domain->usizes[1], (int)domain->usizes[2]};
    double lo[3] = {p0.x, p0.y, p0.z}, hi[3] = {p1.x, p1.y, p1.z};
    bbx_grid_build(res, lo, hi, &cfg.grid);
    if(bbx_create(&cfg, &g_bbx) != BBX_OK){ BbxReport("bbx_create"); return; }
    // vec3f is 3 contiguous Floats (double in the reference's default build): uploads directly as BBX_F64
    static_assert(sizeof(vec3f) == 3 * sizeof(Float), "vec3f must be three packed Floats");
    const int dtype = sizeof(Float) == 8 ? BBX_F64 : BBX_F32;
    if(bbx_set_particles(g_bbx, ps->GetParticleCount(), ps->positions.data, ps->velocities.data, dtype) != BBX_OK)
        BbxReport("bbx_set_particles");
}

// PciSphSolver3::SetColliders: Collider3 + Shape -> bbx_collider (box, sphere, baked SDF, mesh + its SDF grid)
void PciSphSolver3::SetColliders(ColliderSet3 *colliders){
    solverData->sphData->collider = colliders;
    std::vector<bbx_collider> cs(colliders->nColiders);
    std::vector<std::vector<double>> points(colliders->nColiders);
    std::vector<std::vector<int>> tris(colliders->nColiders);
    for(int i = 0; i < colliders->nColiders; i++){
        Collider3 *c = colliders->colliders[i]; Shape *s = c->shape; bbx_collider &b = cs[i];
        memset(&b, 0, sizeof(b));
        b.type = s->type == ShapeSphere ? BBX_COLLIDER_SPHERE : (s->type == ShapeBox ? BBX_COLLIDER_BOX :
                 (s->type == ShapeMesh ? BBX_COLLIDER_MESH : BBX_COLLIDER_SDF));
        b.reverse_orientation = s->reverseOrientation; b.active = c->isActive; b.friction = c->frictionCoefficient;
        BbxRowMajor(s->ObjectToWorld, b.object_to_world, b.world_to_object);
        b.size[0] = s->sizex; b.size[1] = s->sizey; b.size[2] = s->sizez; b.radius = s->radius;
        for(int k = 0; k < 3; k++){ b.linear_velocity[k] = s->linearVelocity[k]; b.angular_velocity[k] = s->angularVelocity[k]; }
        if(s->grid){                       // baked SDF (MakeSDFShape, or GenerateShapeSDF for a mesh): vertex-centred FieldGrid3f
            FieldGrid3f *g = s->grid;      // src/core/grid.h:953-970
            for(int k = 0; k < 3; k++){ b.sdf_resolution[k] = (int)g->resolution[k]; b.sdf_spacing[k] = g->spacing[k]; b.sdf_origin[k] = g->minPoint[k]; }
            static_assert(sizeof(Float) == sizeof(double), "bbx_collider::sdf_field is double: convert when Float is float");
            b.sdf_field = (const double *)g->field;
        }
        if(s->type == ShapeMesh && s->mesh){   // world-space triangles (Transform::Mesh already applied, src/shapes/bvh.cpp:51-56)
            ParsedMesh *m = s->mesh;
            points[i].resize(3 * (size_t)m->nVertices); tris[i].resize(3 * (size_t)m->nTriangles);
            for(int v = 0; v < m->nVertices; v++) for(int k = 0; k < 3; k++) points[i][3 * (size_t)v + k] = m->p[v][k];
            for(int t = 0; t < 3 * m->nTriangles; t++) tris[i][t] = m->indices[t].x;   // Point3i(vertex, normal, uv) per corner
            b.mesh_vertices = m->nVertices; b.mesh_triangles = m->nTriangles;
            b.mesh_points = points[i].data(); b.mesh_indices = tris[i].data();
        }
    }
    if(bbx_set_colliders(g_bbx, (int)cs.size(), cs.data()) != BBX_OK) BbxReport("bbx_set_colliders");
}

// live setter: the reference reads solverData->sphData->viscosity every step
void PciSphSolver3::SetViscosityCoefficient(Float viscosityCoefficient){
    solverData->sphData->viscosity = Max(0, viscosityCoefficient);
    if(g_bbx) bbx_set_param(g_bbx, BBX_PARAM_VISCOSITY, solverData->sphData->viscosity);
}

// PciSphSolver3::Advance (src/solvers/pcisph_solver3.cpp:67-89): CFL sub-stepping + the sub-steps themselves
void PciSphSolver3::Advance(Float timeIntervalInSeconds){
    // colliders a scene script moved or toggled since the last frame (Shape::Update / SetVelocities, ColliderSet3::SetActive)
    ColliderSet3 *cset = solverData->sphData->collider;
    for(int i = 0; cset && i < cset->nColiders; i++){
        Collider3 *c = cset->colliders[i]; Shape *s = c->shape;
        bbx_collider b; memset(&b, 0, sizeof(b));
        b.type = s->type == ShapeSphere ? BBX_COLLIDER_SPHERE : (s->type == ShapeBox ? BBX_COLLIDER_BOX :
                 (s->type == ShapeMesh ? BBX_COLLIDER_MESH : BBX_COLLIDER_SDF));
        if(b.type == BBX_COLLIDER_SPHERE || b.type == BBX_COLLIDER_BOX){   // (grids and meshes are re-sent by SetColliders)
            b.reverse_orientation = s->reverseOrientation; b.active = c->isActive; b.friction = c->frictionCoefficient;
            BbxRowMajor(s->ObjectToWorld, b.object_to_world, b.world_to_object);
            b.size[0] = s->sizex; b.size[1] = s->sizey; b.size[2] = s->sizez; b.radius = s->radius;
            for(int k = 0; k < 3; k++){ b.linear_velocity[k] = s->linearVelocity[k]; b.angular_velocity[k] = s->angularVelocity[k]; }
            bbx_update_collider(g_bbx, i, &b);
        }
        bbx_set_collider_active(g_bbx, i, c->isActive ? 1 : 0);
    }
    int substeps = 0; float ms = 0;
    if(bbx_advance(g_bbx, timeIntervalInSeconds, BBX_SOLVER_PCISPH, &substeps, &ms) != BBX_OK) BbxReport("bbx_advance");
    // the run loop and the serializer read positions (and v, rho) on the host in particle-id order
    ParticleSet3 *ps = solverData->sphData->sphpSet->GetParticleSet();
    const int dtype = sizeof(Float) == 8 ? BBX_F64 : BBX_F32;
    bbx_download(g_bbx, BBX_POSITION, ps->positions.data, dtype);
    bbx_download(g_bbx, BBX_VELOCITY, ps->velocities.data, dtype);
    bbx_download(g_bbx, BBX_DENSITY, ps->densities.data, dtype);
    stepInterval = ms;
}
