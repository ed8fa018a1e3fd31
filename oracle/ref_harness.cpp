// TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Driver for the UNMODIFIED reference (felpzOliveira/Bubbles) CPU path. Compiled by
// oracle/build_ref.sh against the reference sources where they lie (headers by -I, objects from
// /root/reference/src) -- nothing from the reference is copied into this repository.
// Recipe: SURVEY.md Appendix D.  It (a) supplies a calloc shim for the managed-memory arena
// (reference: src/cuda/memory.cpp:58-105 aborts without a CUDA driver), (b) re-runs the body of
// BuildNeighborListKernel on the host (reference: src/core/grid.h:602-624; the launch silently fails
// without a device), (c) runs AdvanceTimeStep (reference: src/solvers/pcisph_solver3.cpp:42-65) or the
// same call sequence phase by phase, dumping every intermediate array as .npy.
//
// Usage: bbref <job.txt>     (commands documented in oracle/README.md)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <fstream>
#include <sstream>
#include <iostream>
#include <chrono>
#include <serializer.h>

#include <pcisph_solver.h>
#include <sph_solver.h>
#include <emitter.h>
#include <collider.h>
#include <shape.h>
#include <grid.h>
#include <util.h>
#include <memory.h>
#include <transform_sequence.h>
#include <obj_loader.h>

// BBREF_GPU: the same driver for the reference's GPU path (its native mode): the reference's own managed-memory
// arena (src/cuda/memory.cpp) is linked instead of the shim, the system stays in GPU mode and every kernel of the
// reference runs on the device with the launch strategy it ships with (16-thread blocks).  Only meaningful on a
// box with a GPU: bench.py times it beside the CPU path ("reference_gpu").
#ifndef BBREF_GPU
// ---- shim for src/cuda/memory.cpp (signatures: src/cuda/cutil.h:134-136, src/cuda/memory.h:17-32)
void *_cudaAllocate(size_t bytes, int, const char *, bool){ return calloc(1, bytes ? bytes : 1); }
void *_cudaAllocateUnregister(size_t bytes, int, const char *, bool){ return calloc(1, bytes ? bytes : 1); }
void *_cudaAllocateExclusive(size_t bytes, int, const char *, bool){ return calloc(1, bytes ? bytes : 1); }
void CudaMemoryManagerStart(const char *){}
std::string CudaGetCurrentKey(){ return std::string("oracle"); }
void CudaMemoryManagerClearCurrent(){}
void CudaMemoryManagerClearAll(){}
#endif
// lives in the (excluded) 2D solver: src/solvers/pcisph_solver2.cpp:7
// (src/core/shape.cpp:411-431, not declared in a header)
bb_cpu_gpu void SetNodeSDFKernel(FieldGrid3f *grid, Shape *shape, int i);
extern const Float kDefaultTimeStepLimitScale = 5.0;

// free functions defined (non-static) in the reference translation units
void AdvanceTimeStep(PciSphSolver3 *solver, Float timeStep, int use_cpu);
void AdvanceTimeStep(SphSolver3 *solver, Float timeStep, int use_cpu);
void PredictVelocityAndPositionCPU(PciSphSolverData3 *data, Float dt, int is_first);
void PredictPressureCPU(PciSphSolverData3 *data, Float delta);
void PredictPressureForceCPU(PciSphSolverData3 *data);
void AccumulateAndIntegrateCPU(PciSphSolverData3 *data, Float timeStep);

// ---------------------------------------------------------------- npy writer
template<typename T> struct NpyType;
template<> struct NpyType<double>{ static const char *s(){ return "<f8"; } };
template<> struct NpyType<int32_t>{ static const char *s(){ return "<i4"; } };
template<> struct NpyType<int64_t>{ static const char *s(){ return "<i8"; } };

template<typename T>
static void WriteNpy(const std::string &path, const T *data, size_t rows, size_t cols){
    FILE *fp = fopen(path.c_str(), "wb");
    if(!fp){ fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
    std::stringstream ss;
    ss << "{'descr': '" << NpyType<T>::s() << "', 'fortran_order': False, 'shape': (" << rows;
    if(cols > 0) ss << ", " << cols << "), }"; else ss << ",), }";
    std::string h = ss.str();
    size_t total = 10 + h.size() + 1;
    size_t pad = (64 - total % 64) % 64;
    h += std::string(pad, ' ');
    h += "\n";
    unsigned char magic[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, 0, 0};
    magic[8] = (unsigned char)(h.size() & 0xff);
    magic[9] = (unsigned char)((h.size() >> 8) & 0xff);
    fwrite(magic, 1, 10, fp);
    fwrite(h.data(), 1, h.size(), fp);
    fwrite(data, sizeof(T), rows * (cols ? cols : 1), fp);
    fclose(fp);
}

static void DumpVec3(const std::string &path, vec3f *v, int n){
    std::vector<double> buf(3 * (size_t)n);
    for(int i = 0; i < n; i++){ buf[3*i] = v[i].x; buf[3*i+1] = v[i].y; buf[3*i+2] = v[i].z; }
    WriteNpy<double>(path, buf.data(), n, 3);
}
static void DumpScalar(const std::string &path, Float *v, int n){
    WriteNpy<double>(path, v, n, 0);
}

// ---------------------------------------------------------------- state
struct Harness{
    Float spacing = 0.02, scale = 1.8, density = WaterDensity;
    Bounds3f domain;
    bool hasDomain = false;
    std::vector<Shape *> shapes;
    std::vector<Float> frictions;
    TransformSequence tseq;          // `tseq_add` / `tseq_restore` / `tseq_eval`
    QuaternionSequence qseq;         // `qseq_add` / `qseq_eval`
    ParticleSetBuilder3 builder;
    ContinuousParticleSetBuilder3 *cbuilder = nullptr; // `continuous <max>`: reserve room so that `append` can add particles between steps
    Grid3 *grid = nullptr;
    PciSphSolver3 pci;
    SphSolver3 sph;
    SphSolverData3 *data = nullptr;
    SphParticleSet3 *sphSet = nullptr;
    ColliderSet3 *colliders = nullptr;
    int solverKind = 0; // 0 = pcisph, 1 = sph
    Float viscosity = -1, drag = -1;
    int gravity = 1;
};

static Transform ReadTransform(std::istringstream &in){
    Float m[4][4];
    for(int i = 0; i < 4; i++) for(int j = 0; j < 4; j++) in >> m[i][j];
    return Transform(m);
}

static void HostBuildNeighborLists(Grid3 *grid){
    // body of BuildNeighborListKernel (src/core/grid.h:602-624) run on the host
    for(unsigned int i = 0; i < grid->total; i++){
        int neighbor[27];
        Cell3 *cell = &grid->cells[i];
        vec3ui u = grid->GetCellIndex(i);
        vec3f center;
        for(int k = 0; k < 3; k++)
            center[k] = grid->minPoint[k] + u[k] * grid->cellsLen[k] + 0.5 * grid->cellsLen[k];
        vec3f pMin = center - 0.5 * grid->cellsLen;
        vec3f pMax = center + 0.5 * grid->cellsLen;
        int count = grid->GetNeighborListFor(i, 1, &neighbor[0]);
        cell->SetNeighborListPtr(&grid->neighborListPtr[27 * i]);
        cell->Set(Bounds3f(pMin, pMax), i);
        cell->SetNeighborList(&neighbor[0], count);
    }
}

static void DumpGrid(Harness &H, const std::string &prefix){
    Grid3 *g = H.grid;
    ParticleSet3 *pSet = H.sphSet->GetParticleSet();
    int n = pSet->GetParticleCount();
    std::vector<int32_t> counts(g->total), order;
    order.reserve(n);
    for(unsigned int c = 0; c < g->total; c++){
        Cell3 *cell = &g->cells[c];
        int len = cell->GetChainLength();
        counts[c] = len;
        ParticleChain *p = cell->GetChain();
        for(int j = 0; j < len; j++){ order.push_back((int32_t)p->pId); p = p->next; }
    }
    WriteNpy<int32_t>(prefix + "cell_count.npy", counts.data(), counts.size(), 0);
    WriteNpy<int32_t>(prefix + "cell_order.npy", order.data(), order.size(), 0);
    std::vector<int32_t> bcount(n), bids((size_t)n * MaximumParticlesPerBucket, -1);
    for(int i = 0; i < n; i++){
        Bucket *b = pSet->GetParticleBucket(i);
        bcount[i] = b->Count();
        for(int k = 0; k < b->Count(); k++) bids[(size_t)i * MaximumParticlesPerBucket + k] = b->Get(k);
    }
    WriteNpy<int32_t>(prefix + "nbr_count.npy", bcount.data(), n, 0);
    WriteNpy<int32_t>(prefix + "nbr_ids.npy", bids.data(), n, MaximumParticlesPerBucket);
}

static void DumpState(Harness &H, const std::string &prefix){
    ParticleSet3 *pSet = H.sphSet->GetParticleSet();
    int n = pSet->GetParticleCount();
    DumpVec3(prefix + "pos.npy", pSet->positions.data, n);
    DumpVec3(prefix + "vel.npy", pSet->velocities.data, n);
    DumpVec3(prefix + "force.npy", pSet->forces.data, n);
    DumpScalar(prefix + "density.npy", pSet->densities.data, n);
    DumpScalar(prefix + "pressure.npy", pSet->pressures.data, n);
}

static void Setup(Harness &H){
    if(!H.hasDomain){ fprintf(stderr, "no domain\n"); exit(2); }
    H.grid = UtilBuildGridForDomain(H.domain, H.spacing, H.scale);
#ifndef BBREF_GPU
    HostBuildNeighborLists(H.grid);
#endif
    ColliderSetBuilder3 cBuilder;
    for(size_t i = 0; i < H.shapes.size(); i++) cBuilder.AddCollider3(H.shapes[i], H.frictions[i]);
    H.colliders = cBuilder.GetColliderSet();
    if(H.cbuilder){ H.cbuilder->Commit(); H.sphSet = SphParticleSet3FromContinuousBuilder(H.cbuilder); }
    else H.sphSet = SphParticleSet3FromBuilder(&H.builder);
    H.sphSet->SetRelativeKernelRadius(H.scale);
    H.data = DefaultSphSolverData3(H.gravity != 0);
    if(H.solverKind == 0){
        H.pci.Initialize(H.data);
        H.pci.Setup(H.density, H.spacing, H.scale, H.grid, H.sphSet);
        H.pci.SetColliders(H.colliders);
        if(H.viscosity >= 0) H.pci.SetViscosityCoefficient(H.viscosity);
    }else{
        H.sph.Initialize(H.data);
        H.sph.Setup(H.density, H.spacing, H.scale, H.grid, H.sphSet);
        H.sph.SetColliders(H.colliders);
        if(H.viscosity >= 0) H.sph.SetViscosityCoefficient(H.viscosity);
    }
    if(H.drag >= 0) H.data->dragCoefficient = H.drag;
    if(H.cbuilder){ H.cbuilder->SetKernelRadius(H.spacing * H.scale); H.cbuilder->MapGrid(H.grid); }
    ParticleSet3 *pSet = H.sphSet->GetParticleSet();
    vec3ui res = H.grid->GetIndexCount();
    printf("[bbref] N=%d cells=%u (%u x %u x %u) mass=%.17g h=%.17g\n", pSet->GetParticleCount(),
           H.grid->GetCellCount(), res.x, res.y, res.z, pSet->GetMass(), H.sphSet->GetKernelRadius());
    printf("[bbref] grid min=(%.17g %.17g %.17g) len=(%.17g %.17g %.17g)\n",
           H.grid->minPoint.x, H.grid->minPoint.y, H.grid->minPoint.z,
           H.grid->cellsLen.x, H.grid->cellsLen.y, H.grid->cellsLen.z);
}

// One PCISPH sub-step, phase by phase, same call sequence as
// AdvanceTimeStep (pcisph_solver3.cpp:42-65) + ComputePressureForceAndIntegrate (pcisph_equations3.cpp:218-256)
static void TraceStep(Harness &H, Float dt, const std::string &prefix){
    SphSolverData3 *data = H.data;
    PciSphSolverData3 *pd = H.pci.solverData;
    ParticleSet3 *pSet = H.sphSet->GetParticleSet();
    int n = pSet->GetParticleCount();
    int32_t flag = data->sphpSet->requiresHigherLevelUpdate;
    WriteNpy<int32_t>(prefix + "rebuild_flag.npy", &flag, 1, 0);
    UpdateGridDistributionCPU(data);
    data->sphpSet->ResetHigherLevel();
    DumpGrid(H, prefix);
    ComputeParticleInteractionCPU(data);
    ComputeDensityCPU(data);
    DumpScalar(prefix + "density.npy", pSet->densities.data, n);
    DumpScalar(prefix + "eos_pressure.npy", pSet->pressures.data, n); // ComputePressureValue (Tait EOS) of ComputeDensityFor
    ComputeNonPressureForceCPU(data);
    DumpVec3(prefix + "force_np.npy", pSet->forces.data, n);
    Float delta = H.pci.ComputeDelta(dt);
    WriteNpy<double>(prefix + "delta.npy", &delta, 1, 0);
    // k = 0 only: the reference's loop exits after one iteration
    PredictVelocityAndPositionCPU(pd, dt, 1);
    DumpVec3(prefix + "pos_pred.npy", pd->tempPositions, n);
    PredictPressureCPU(pd, delta);
    DumpScalar(prefix + "density_pred.npy", pd->densityPredicted, n);
    DumpScalar(prefix + "pressure.npy", pSet->pressures.data, n);
    PredictPressureForceCPU(pd);
    DumpVec3(prefix + "force_p.npy", pd->pressureForces, n);
    Float maxErr = 0;
    for(int i = 0; i < n; i++){ Float e = pd->densityErrors[i]; if(e*e > maxErr*maxErr) maxErr = e; }
    WriteNpy<double>(prefix + "max_density_error.npy", &maxErr, 1, 0);
    AccumulateAndIntegrateCPU(pd, dt);
    ComputePseudoViscosityInterpolationCPU(data, dt);
    DumpVec3(prefix + "pos_out.npy", pSet->positions.data, n);
    DumpVec3(prefix + "vel_out.npy", pSet->velocities.data, n);
    DumpVec3(prefix + "force_out.npy", pSet->forces.data, n);
    int32_t flag2 = data->sphpSet->requiresHigherLevelUpdate;
    WriteNpy<int32_t>(prefix + "rebuild_flag_out.npy", &flag2, 1, 0);
}

int main(int argc, char **argv){
    if(argc < 2){ fprintf(stderr, "usage: bbref job.txt\n"); return 2; }
    cudaSetLaunchStrategy(CudaLaunchStrategy::CustomizedBlockSize, 16);
#ifndef BBREF_GPU
    SetSystemUseCPU();
    SetCPUThreads(1);
#else
    { int nd = 0; if(cudaGetDeviceCount(&nd) != cudaSuccess || nd == 0){ fprintf(stderr, "bbref_gpu: no CUDA device\n"); return 3; } }
#endif
    Harness H;
    std::ifstream job(argv[1]);
    if(!job){ fprintf(stderr, "cannot open job %s\n", argv[1]); return 2; }
    std::string line;
    while(std::getline(job, line)){
        if(line.empty() || line[0] == '#') continue;
        std::istringstream in(line);
        std::string cmd; in >> cmd;
        if(cmd == "threads"){ int t; in >> t; SetCPUThreads(t); }
        else if(cmd == "solver"){ std::string s; in >> s; H.solverKind = (s == "sph") ? 1 : 0; }
        else if(cmd == "spacing"){ in >> H.spacing; }
        else if(cmd == "scale"){ in >> H.scale; }
        else if(cmd == "density"){ in >> H.density; }
        else if(cmd == "viscosity"){ in >> H.viscosity; }
        else if(cmd == "drag"){ in >> H.drag; }
        else if(cmd == "gravity"){ in >> H.gravity; }
        else if(cmd == "domain"){
            vec3f a, b; in >> a.x >> a.y >> a.z >> b.x >> b.y >> b.z;
            H.domain = Bounds3f(a, b); H.hasDomain = true;
        }
        else if(cmd == "domain_from_collider"){
            int idx; in >> idx; H.domain = H.shapes[idx]->GetBounds(); H.hasDomain = true;
        }
        else if(cmd == "collider"){
            std::string kind; in >> kind;
            if(kind == "box"){
                Transform t = ReadTransform(in);
                vec3f size; int rev; Float fr;
                in >> size.x >> size.y >> size.z >> rev >> fr;
                H.shapes.push_back(MakeBox(t, size, rev != 0)); H.frictions.push_back(fr);
            }else if(kind == "sphere"){
                Transform t = ReadTransform(in);
                Float r; int rev; Float fr;
                in >> r >> rev >> fr;
                H.shapes.push_back(MakeSphere(t, r, rev != 0)); H.frictions.push_back(fr);
            }else if(kind == "sdf"){
                // SDF grid collider from a raw field file written by the test (analytic SDF sampled at
                // the node positions this harness reports); mirrors Shape::InitSDFShape (shape.h:201-230)
                vec3f a, b; Float dx, margin, fr; std::string file;
                in >> a.x >> a.y >> a.z >> b.x >> b.y >> b.z >> dx >> margin >> fr >> file;
                Shape *shape = (Shape *)calloc(1, sizeof(Shape));
                Bounds3f bounds(a, b);
                shape->bounds = bounds;
                shape->type = ShapeType::ShapeSDF;
                shape->WorldToObject = Transform();
                shape->ObjectToWorld = Transform();
                vec3f sc(bounds.ExtentOn(0), bounds.ExtentOn(1), bounds.ExtentOn(2));
                shape->bounds.pMin -= margin * sc;
                shape->bounds.pMax += margin * sc;
                Float width = shape->bounds.ExtentOn(0), height = shape->bounds.ExtentOn(1);
                Float depth = shape->bounds.ExtentOn(2);
                int resolution = (int)std::ceil(width / dx);
                dx = width / (Float)resolution;
                int resolutionY = (int)std::ceil(resolution * height / width);
                int resolutionZ = (int)std::ceil(resolution * depth / width);
                shape->grid = (FieldGrid3f *)calloc(1, sizeof(FieldGrid3f));
                shape->grid->Build(vec3ui(resolution, resolutionY, resolutionZ), vec3f(dx),
                                   shape->bounds.pMin, VertexCentered);
                FILE *fp = fopen(file.c_str(), "rb");
                if(!fp){ fprintf(stderr, "cannot open sdf field %s\n", file.c_str()); return 2; }
                size_t got = fread(shape->grid->field, sizeof(Float), shape->grid->total, fp);
                fclose(fp);
                if(got != shape->grid->total){
                    fprintf(stderr, "sdf field size mismatch: file %zu, grid %u (%u %u %u)\n", got,
                            shape->grid->total, shape->grid->resolution.x, shape->grid->resolution.y,
                            shape->grid->resolution.z);
                    return 2;
                }
                shape->grid->MarkFilled();
                H.shapes.push_back(shape); H.frictions.push_back(fr);
            }else if(kind == "mesh"){
                // triangle-mesh collider: MakeMesh (src/shapes/bvh.cpp:51-56: BVH over the triangles) + the SDF the collider set
                // generates for it, GenerateShapeSDF (src/core/shape.cpp:479-511) -- whose kernel launch is replaced by a host
                // loop over the same per-node function (SetNodeSDFKernel, shape.cpp:411-431: BVH closest distance, sign by
                // ray parity), because there is no device here.  File: int64 nv, nt; nv x 3 doubles; nt x 3 int32.
                std::string file; int rev; Float fr, dx, margin;
                in >> file >> rev >> fr >> dx >> margin;
                FILE *fp = fopen(file.c_str(), "rb");
                if(!fp){ fprintf(stderr, "cannot open mesh %s\n", file.c_str()); return 2; }
                int64_t nv = 0, nt = 0; size_t got = fread(&nv, 8, 1, fp) + fread(&nt, 8, 1, fp);
                std::vector<double> vp(3 * nv); std::vector<int32_t> vi(3 * nt);
                got += fread(vp.data(), 8, 3 * nv, fp) + fread(vi.data(), 4, 3 * nt, fp);
                fclose(fp);
                if(got != (size_t)(2 + 3 * nv + 3 * nt)){ fprintf(stderr, "short mesh file\n"); return 2; }
                ParsedMesh *mesh = (ParsedMesh *)calloc(1, sizeof(ParsedMesh));
                mesh->p = (Point3f *)calloc(nv, sizeof(Point3f));
                mesh->indices = (Point3i *)calloc(3 * nt, sizeof(Point3i));
                for(int64_t i = 0; i < nv; i++) mesh->p[i] = Point3f(vp[3 * i], vp[3 * i + 1], vp[3 * i + 2]);
                for(int64_t i = 0; i < 3 * nt; i++) mesh->indices[i] = Point3i(vi[i], 0, 0);
                mesh->nTriangles = (int)nt; mesh->nVertices = (int)nv;
                snprintf(mesh->name, sizeof(mesh->name), "harness-mesh");
                Shape *shape = MakeMesh(mesh, Transform(), rev != 0);
                {   // GenerateShapeSDF(shape, dx, margin), host loop instead of GPULaunch(CreateShapeSDFGPU)
                    Bounds3f bounds = shape->GetBounds();
                    vec3f sc(bounds.ExtentOn(0), bounds.ExtentOn(1), bounds.ExtentOn(2));
                    bounds.pMin -= margin * sc; bounds.pMax += margin * sc;
                    Float width = bounds.ExtentOn(0), height = bounds.ExtentOn(1), depth = bounds.ExtentOn(2);
                    int resolution = (int)std::ceil(width / dx);
                    dx = width / (Float)resolution;
                    int resolutionY = (int)std::ceil(resolution * height / width);
                    int resolutionZ = (int)std::ceil(resolution * depth / width);
                    shape->grid = (FieldGrid3f *)calloc(1, sizeof(FieldGrid3f));
                    shape->grid->Build(vec3ui(resolution, resolutionY, resolutionZ), vec3f(dx), bounds.pMin, VertexCentered);
                    for(unsigned int i = 0; i < shape->grid->total; i++) SetNodeSDFKernel(shape->grid, shape, (int)i);
                    shape->grid->MarkFilled();
                }
                H.shapes.push_back(shape); H.frictions.push_back(fr);
            }else{ fprintf(stderr, "unknown collider %s\n", kind.c_str()); return 2; }
        }
        else if(cmd == "emit_box"){
            // reference emitter path (emitter.cpp:242-330): BCC lattice + libc rand() jitter
            Transform t = ReadTransform(in);
            vec3f size, v; Float jitter; unsigned seed;
            in >> size.x >> size.y >> size.z >> v.x >> v.y >> v.z >> jitter >> seed;
            srand(seed);
            Shape *box = MakeBox(t, size);
            VolumeParticleEmitterSet3 set;
            VolumeParticleEmitter3 em(box, box->GetBounds(), H.spacing, v);
            set.AddEmitter(&em);
            set.SetJitter(jitter);
            set.Emit(&H.builder);
        }
        else if(cmd == "particles"){
            std::string file; in >> file;
            FILE *fp = fopen(file.c_str(), "rb");
            if(!fp){ fprintf(stderr, "cannot open %s\n", file.c_str()); return 2; }
            int64_t n = 0;
            if(fread(&n, sizeof(n), 1, fp) != 1) return 2;
            std::vector<double> pos(3 * n), vel(3 * n);
            if(fread(pos.data(), sizeof(double), 3 * n, fp) != (size_t)(3 * n)) return 2;
            if(fread(vel.data(), sizeof(double), 3 * n, fp) != (size_t)(3 * n)) return 2;
            fclose(fp);
            for(int64_t i = 0; i < n; i++){
                vec3f p(pos[3*i], pos[3*i+1], pos[3*i+2]), v(vel[3*i], vel[3*i+1], vel[3*i+2]);
                if(H.cbuilder) H.cbuilder->AddParticle(p, v); else H.builder.AddParticle(p, v);
            }
        }
        else if(cmd == "pseudo"){
            // pseudo-viscosity coefficient (SphSolverData3::pseudoViscosity, default 10: the smoothing only runs when
            // coefficient * dt > 0.1, sph_equations3.cpp:469-483); after `setup`
            Float c; in >> c; if(!H.data){ fprintf(stderr, "pseudo: after setup\n"); return 2; } H.data->pseudoViscosity = c;
        }
        else if(cmd == "continuous"){ int maxp; in >> maxp; H.cbuilder = new ContinuousParticleSetBuilder3(maxp); }
        else if(cmd == "map_emit"){
            // ContinuousParticleSetBuilder3::MapGridEmit (src/core/grid.h:1367-1407): re-emit the particles of the cells
            // mapped at setup wherever the cell has room and no particle closer than d; constant emission velocity
            Float vx, vy, vz, d; in >> vx >> vy >> vz >> d;
            if(!H.cbuilder){ fprintf(stderr, "map_emit needs `continuous <max>` before the particles\n"); return 2; }
            int before = H.sphSet->GetParticleSet()->GetParticleCount();
            H.cbuilder->MapGridEmit([&](const vec3f &) -> vec3f { return vec3f(vx, vy, vz); }, d);
            printf("[bbref] map_emit added=%d total=%d\n", H.sphSet->GetParticleSet()->GetParticleCount() - before,
                   H.sphSet->GetParticleSet()->GetParticleCount());
        }
        else if(cmd == "append"){
            // ContinuousParticleSetBuilder3::AddParticle + Commit (src/core/grid.h:1409-1441): AppendData, then
            // DistributeByParticleList puts the new ids at the tail of their cells' chains
            std::string file; in >> file;
            if(!H.cbuilder){ fprintf(stderr, "append needs `continuous <max>` before the particles\n"); return 2; }
            FILE *fp = fopen(file.c_str(), "rb");
            if(!fp){ fprintf(stderr, "cannot open %s\n", file.c_str()); return 2; }
            int64_t n = 0;
            if(fread(&n, sizeof(n), 1, fp) != 1) return 2;
            std::vector<double> pos(3 * n), vel(3 * n);
            if(fread(pos.data(), sizeof(double), 3 * n, fp) != (size_t)(3 * n)) return 2;
            if(fread(vel.data(), sizeof(double), 3 * n, fp) != (size_t)(3 * n)) return 2;
            fclose(fp);
            for(int64_t i = 0; i < n; i++)
                if(!H.cbuilder->AddParticle(vec3f(pos[3*i], pos[3*i+1], pos[3*i+2]), vec3f(vel[3*i], vel[3*i+1], vel[3*i+2]))){ fprintf(stderr, "append: builder full\n"); return 2; }
            H.cbuilder->Commit();
            printf("[bbref] append n=%lld total=%d\n", (long long)n, H.sphSet->GetParticleSet()->GetParticleCount());
        }
        else if(cmd == "setup"){ Setup(H); }
        else if(cmd == "set_chains"){
            // inject an explicit per-cell chain order (counts + concatenated ids), built with the
            // reference's own Cell::AddToChain (grid.h:128-141) after a reset
            std::string fc, fo; in >> fc >> fo;
            Grid3 *g = H.grid; ParticleSet3 *pSet = H.sphSet->GetParticleSet();
            std::vector<int32_t> counts(g->total), order(pSet->GetParticleCount());
            FILE *fp = fopen(fc.c_str(), "rb"); size_t r = fread(counts.data(), 4, counts.size(), fp); fclose(fp);
            fp = fopen(fo.c_str(), "rb"); r += fread(order.data(), 4, order.size(), fp); fclose(fp);
            size_t at = 0;
            for(unsigned int c = 0; c < g->total; c++){
                g->DistributeResetCell(c);
                for(int k = 0; k < counts[c]; k++){
                    int pid = order[at++];
                    ParticleChain *node = pSet->GetParticleChainNode(pid);
                    node->cId = c; node->pId = pid; node->sId = pSet->GetFamilyId();
                    g->cells[c].AddToChain(node);
                }
            }
            H.sphSet->ResetHigherLevel();
        }
        else if(cmd == "round32"){
            ParticleSet3 *pSet = H.sphSet->GetParticleSet();
            for(int i = 0; i < pSet->GetParticleCount(); i++){
                vec3f p = pSet->positions.data[i], v = pSet->velocities.data[i];
                pSet->positions.data[i] = vec3f((float)p.x, (float)p.y, (float)p.z);
                pSet->velocities.data[i] = vec3f((float)v.x, (float)v.y, (float)v.z);
            }
        }
        else if(cmd == "step"){
            Float dt; int n; in >> dt >> n;
            auto t0 = std::chrono::steady_clock::now();
            for(int s = 0; s < n; s++){
                if(H.solverKind == 0) AdvanceTimeStep(&H.pci, dt, 1);
                else AdvanceTimeStep(&H.sph, dt, 1);
            }
#ifdef BBREF_GPU
            cudaDeviceSynchronize();
#endif
            double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            int np = H.sphSet->GetParticleSet()->GetParticleCount();
            printf("[bbref] steps=%d dt=%g seconds=%.6f particle_updates_per_s=%.6e\n", n, (double)dt, sec,
                   (double)np * n / sec);
        }
        else if(cmd == "advance"){
            Float t; in >> t;
            if(H.solverKind == 0) H.pci.Advance(t); else H.sph.Advance(t);
        }
        else if(cmd == "trace"){ Float dt; std::string prefix; in >> dt >> prefix; TraceStep(H, dt, prefix); }
        else if(cmd == "dump"){ std::string prefix; in >> prefix; DumpState(H, prefix); }
        else if(cmd == "dump_grid"){ std::string prefix; in >> prefix; DumpGrid(H, prefix); }
        else if(cmd == "delta"){
            Float dt; in >> dt;
            printf("[bbref] delta(%.17g)=%.17g\n", (double)dt, (double)H.pci.ComputeDelta(dt));
        }
        else if(cmd == "collide"){
            // apply ColliderSet3::ResolveCollision to a list of (pos, vel): file in, npy out
            std::string file, prefix; Float radius, rest; in >> file >> radius >> rest >> prefix;
            FILE *fp = fopen(file.c_str(), "rb");
            int64_t n = 0; size_t r = fread(&n, sizeof(n), 1, fp);
            std::vector<double> pos(3 * n), vel(3 * n);
            r += fread(pos.data(), 8, 3 * n, fp); r += fread(vel.data(), 8, 3 * n, fp); fclose(fp);
            std::vector<int32_t> hit(n);
            for(int64_t i = 0; i < n; i++){
                vec3f p(pos[3*i], pos[3*i+1], pos[3*i+2]), v(vel[3*i], vel[3*i+1], vel[3*i+2]);
                hit[i] = H.colliders->ResolveCollision(radius, rest, &p, &v) ? 1 : 0;
                pos[3*i] = p.x; pos[3*i+1] = p.y; pos[3*i+2] = p.z;
                vel[3*i] = v.x; vel[3*i+1] = v.y; vel[3*i+2] = v.z;
            }
            WriteNpy<double>(prefix + "pos.npy", pos.data(), n, 3);
            WriteNpy<double>(prefix + "vel.npy", vel.data(), n, 3);
            WriteNpy<int32_t>(prefix + "hit.npy", hit.data(), n, 0);
        }
        else if(cmd == "dump_sdf"){
            // the SDF grid attached to shape idx: node counts, spacing, origin (position of node 0), field (x fastest)
            int idx; std::string prefix; in >> idx >> prefix;
            FieldGrid3f *gr = H.shapes[idx]->grid;
            if(!gr){ fprintf(stderr, "shape %d has no sdf grid\n", idx); return 2; }
            int32_t res[3] = {(int32_t)gr->resolution.x, (int32_t)gr->resolution.y, (int32_t)gr->resolution.z};
            vec3f p0 = gr->GetDataPosition(vec3ui(0, 0, 0));
            double meta[6] = {(double)gr->spacing.x, (double)gr->spacing.y, (double)gr->spacing.z, (double)p0.x, (double)p0.y, (double)p0.z};
            WriteNpy<int32_t>(prefix + "res.npy", res, 3, 0);
            WriteNpy<double>(prefix + "meta.npy", meta, 6, 0);
            WriteNpy<double>(prefix + "field.npy", gr->field, gr->total, 0);
            Bounds3f b = H.shapes[idx]->GetBounds();
            double bb[6] = {(double)b.pMin.x, (double)b.pMin.y, (double)b.pMin.z, (double)b.pMax.x, (double)b.pMax.y, (double)b.pMax.z};
            WriteNpy<double>(prefix + "bounds.npy", bb, 6, 0);
        }
        else if(cmd == "closest_distance"){
            // Shape::ClosestDistance of shape idx at a list of points (mesh: BVHMeshClosestDistance, bvh.cpp:500-557)
            int idx; std::string file, prefix; in >> idx >> file >> prefix;
            FILE *fp = fopen(file.c_str(), "rb");
            int64_t n = 0; size_t r = fread(&n, sizeof(n), 1, fp);
            std::vector<double> pos(3 * n), out(n);
            r += fread(pos.data(), 8, 3 * n, fp); fclose(fp);
            for(int64_t i = 0; i < n; i++) out[i] = H.shapes[idx]->ClosestDistance(vec3f(pos[3*i], pos[3*i+1], pos[3*i+2]));
            WriteNpy<double>(prefix + "distance.npy", out.data(), n, 0);
        }
        else if(cmd == "save_frame"){
            // the reference's own frame writer (SerializerSaveSphDataSet3, src/third/serializer.cpp:884-921) on the current state
            std::string file; int flags; in >> file >> flags;
            SerializerSaveSphDataSet3(H.data, file.c_str(), flags);
        }
        else if(cmd == "save_sim"){
            // UtilSaveSimulation3 (src/core/util.h:296-328): shape blocks of every collider but the last + the particle block
            std::string file; int flags; in >> file >> flags;
            if(H.solverKind == 0) UtilSaveSimulation3<PciSphSolver3, ParticleSet3>(&H.pci, H.sphSet->GetParticleSet(), file.c_str(), flags);
            else UtilSaveSimulation3<SphSolver3, ParticleSet3>(&H.sph, H.sphSet->GetParticleSet(), file.c_str(), flags);
        }
        else if(cmd == "load_frame"){
            // the reference's own frame reader (what bbtool uses: SerializerLoadParticles3, serializer.cpp:444-559)
            std::string file, prefix; in >> file >> prefix;
            std::vector<SerializedParticle> ps; int flags = 0;
            int n = SerializerLoadParticles3(&ps, file.c_str(), flags);
            std::vector<double> pos(3 * (size_t)std::max(n, 0)), vel(3 * (size_t)std::max(n, 0)), rho((size_t)std::max(n, 0)), mass((size_t)std::max(n, 0));
            for(int i = 0; i < n; i++){
                pos[3*i] = ps[i].position.x; pos[3*i+1] = ps[i].position.y; pos[3*i+2] = ps[i].position.z;
                vel[3*i] = ps[i].velocity.x; vel[3*i+1] = ps[i].velocity.y; vel[3*i+2] = ps[i].velocity.z;
                rho[i] = ps[i].density; mass[i] = (flags & SERIALIZER_MASS) ? ps[i].mass : 0.0;
            }
            WriteNpy<double>(prefix + "pos.npy", pos.data(), n, 3);
            WriteNpy<double>(prefix + "vel.npy", vel.data(), n, 3);
            WriteNpy<double>(prefix + "rho.npy", rho.data(), n, 0);
            WriteNpy<double>(prefix + "mass.npy", mass.data(), n, 0);
            printf("[bbref] load_frame count=%d flags=%d\n", n, flags);
        }
        else if(cmd == "collider_velocity"){
            // Shape::SetVelocities (src/core/shape.cpp:290-298): rigid motion of collider idx (only the sphere's closest-point
            // query carries VelocityAt into the response)
            int idx; vec3f v, w; in >> idx >> v.x >> v.y >> v.z >> w.x >> w.y >> w.z;
            H.shapes[idx]->SetVelocities(v, w);
        }
        else if(cmd == "collider_active"){
            // ColliderSet3::SetActive (src/core/collider.cpp:232-236), after setup
            int idx, on; in >> idx >> on; H.colliders->SetActive(idx, on != 0);
        }
        else if(cmd == "sdf_nodes"){
            // report the node layout of SDF collider idx so the test can sample its analytic SDF there
            int idx; in >> idx; FieldGrid3f *g = H.shapes[idx]->grid;
            printf("[bbref] sdf res=%u %u %u spacing=%.17g origin=%.17g %.17g %.17g\n", g->resolution.x,
                   g->resolution.y, g->resolution.z, g->spacing.x, g->minPoint.x, g->minPoint.y, g->minPoint.z);
        }
        else if(cmd == "load_obj"){
            // the reference's .obj loader (src/third/obj_loader.cpp:434-625): vertices in first-use order and the vertex index
            // of every triangle corner (mesh->indices[].x), as .npy
            std::string file, prefix; in >> file >> prefix;
            ParsedMesh *m = LoadObj(file.c_str());
            std::vector<double> pts(3 * (size_t)m->nVertices); std::vector<int32_t> tri(3 * (size_t)m->nTriangles);
            for(int i = 0; i < m->nVertices; i++) for(int k = 0; k < 3; k++) pts[3 * (size_t)i + k] = m->p[i][k];
            for(int i = 0; i < 3 * m->nTriangles; i++) tri[(size_t)i] = m->indices[i].x;
            WriteNpy<double>(prefix + "points.npy", pts.data(), m->nVertices, 3);
            WriteNpy<int32_t>(prefix + "triangles.npy", tri.data(), m->nTriangles, 3);
        }
        else if(cmd == "set_boundary"){
            // per-particle boundary layer (ParticleSet3::v0s, what the boundary classifiers of src/boundaries/* leave there) and
            // normals, from files: int64 n, double v0[n] | int64 n, double normal[3n]
            std::string fb, fn; in >> fb >> fn;
            ParticleSet3 *ps = H.sphSet->GetParticleSet();
            FILE *fp = fopen(fb.c_str(), "rb"); int64_t n = 0; size_t r = fread(&n, 8, 1, fp);
            std::vector<double> v0(n); r += fread(v0.data(), 8, n, fp); fclose(fp);
            fp = fopen(fn.c_str(), "rb"); r += fread(&n, 8, 1, fp);
            std::vector<double> nr(3 * n); r += fread(nr.data(), 8, 3 * n, fp); fclose(fp);
            for(int64_t i = 0; i < n; i++){ ps->SetParticleV0((int)i, v0[i]); ps->SetParticleNormal((int)i, vec3f(nr[3*i], nr[3*i+1], nr[3*i+2])); }
        }
        else if(cmd == "save_frame_b"){
            // SerializerSaveSphDataSet3 with the boundary vector UtilGetBoundaryState builds (src/core/util.h:708-725)
            std::string file; int flags; in >> file >> flags;
            std::vector<int> boundaries;
            UtilGetBoundaryState(H.sphSet->GetParticleSet(), &boundaries);
            SerializerSaveSphDataSet3(H.data, file.c_str(), flags, &boundaries);
        }
        else if(cmd == "tseq_add"){
            // TransformSequence::AddInterpolation(&K0, &K1, s0, s1) (src/core/transform_sequence.cpp:30-47) between two
            // keyframes K = Translate(t) * Rotate(angle [deg], axis) * Scale(s)
            Float a[16]; for(int k = 0; k < 16; k++) in >> a[k];
            Float s0, s1; in >> s0 >> s1;
            Transform k0 = Translate(vec3f(a[0], a[1], a[2])) * Rotate(a[3], vec3f(a[4], a[5], a[6])) * Scale(a[7]);
            Transform k1 = Translate(vec3f(a[8], a[9], a[10])) * Rotate(a[11], vec3f(a[12], a[13], a[14])) * Scale(a[15]);
            H.tseq.AddInterpolation(&k0, &k1, s0, s1);
        }
        else if(cmd == "tseq_restore"){ Float s0, s1; in >> s0 >> s1; H.tseq.AddRestore(s0, s1); }
        else if(cmd == "tseq_eval" || cmd == "qseq_eval"){
            // Interpolate(t, &transform, &linear, &angular) at t = t0 + i dt, i < n, in call order (the sequence remembers
            // its last result: transform_sequence.cpp:103-160)
            Float t0, dt; int n; std::string prefix; in >> t0 >> dt >> n >> prefix;
            std::vector<double> m(16 * (size_t)n), mi(16 * (size_t)n), lin(3 * (size_t)n), ang(3 * (size_t)n);
            for(int i = 0; i < n; i++){
                Transform tr; vec3f l(0), w(0);
                if(cmd == "tseq_eval") H.tseq.Interpolate(t0 + i * dt, &tr, &l, &w);
                else H.qseq.Interpolate(t0 + i * dt, &tr, &w);
                for(int r = 0; r < 4; r++) for(int c = 0; c < 4; c++){ m[16 * (size_t)i + 4 * r + c] = tr.m.m[r][c]; mi[16 * (size_t)i + 4 * r + c] = tr.mInv.m[r][c]; }
                for(int k = 0; k < 3; k++){ lin[3 * (size_t)i + k] = l[k]; ang[3 * (size_t)i + k] = w[k]; }
            }
            WriteNpy<double>(prefix + "m.npy", m.data(), n, 16);
            WriteNpy<double>(prefix + "minv.npy", mi.data(), n, 16);
            WriteNpy<double>(prefix + "linear.npy", lin.data(), n, 3);
            WriteNpy<double>(prefix + "angular.npy", ang.data(), n, 3);
        }
        else if(cmd == "qseq_add"){
            // QuaternionSequence::AddQuaternion(angle [deg], axis, t) (transform_sequence.cpp:170-177)
            Float ang, x, y, z, t; in >> ang >> x >> y >> z >> t;
            H.qseq.AddQuaternion(ang, vec3f(x, y, z), t);
        }
        else{ fprintf(stderr, "unknown command: %s\n", cmd.c_str()); return 2; }
    }
    return 0;
}
