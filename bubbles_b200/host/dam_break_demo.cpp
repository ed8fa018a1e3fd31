// dam_break_demo -- the reference's test_pcisph3_dam_break scene (src/tests/test_pcisph_extra.cpp:1102-1169)
// written against the bbx C++ facade (bubbles_api.h): same constants, same call sequence
// (MakeBox / VolumeParticleEmitter3 / UtilBuildGridForDomain / ColliderSetBuilder3 / PciSphSolver3 /
// SerializerSaveSphDataSet3 / PciSphRunSimulation3), host code only -- every kernel runs inside libbbx.so.
//
//   dam_break_demo [--scaling S] [--frames N] [--steps N --dt X] [--jitter J] [--out DIR] [--dump FILE] [--sph] [--emit K] [--map-emit] [--load FRAME] [--keyframes] [--mesh]
//     --scaling  domainScaling of the reference scene (2.5 there: ~0.9 M particles; default 0.6: ~12 k)
//     --frames   frames of 1/240 s through Advance() (CFL sub-stepping), default 2
//     --steps    instead of frames: N fixed-dt sub-steps (AdvanceTimeStep)
//     --out      directory for the text frames bbtool reads (out_<frame>.txt), off by default
//     --emit     continuous emission (ContinuousParticleSetBuilder3): after every frame K more particles enter above the
//                fluid (AddParticle + Commit), as the reference's MapGridEmit scenes do
//     --map-emit continuous RE-emission (ContinuousParticleSetBuilder3::MapGrid after Setup, MapGridEmit after every frame,
//                src/core/grid.h:1288-1407): the cells the block started in are refilled where there is room; the per-cell
//                test runs on the device (bbx_query_cells)
//     --load     start from a frame file (positions and, when present, velocities: SerializerLoadSphDataSet3) instead of
//                emitting the block, e.g. the reference's resources/dam_break_50
//     --keyframes a sphere obstacle on the floor driven by a TransformSequence (three keyframes + AddRestore), updated every
//                frame with Shape::Update / SetVelocities exactly as the reference's moving-container scene does
//                (src/tests/test_pcisph3.cpp:82-136), handed to the engine by PciSphSolver3::UpdateCollider
//     --mesh     a triangle-mesh obstacle (an octahedron standing on the floor in the path of the flow): MakeMesh +
//                ColliderSet3::GenerateSDFs (host bake, mesh_sdf.h) -> BBX_COLLIDER_MESH
//     --dump     raw little-endian doubles: n, then n x 3 positions, n x 3 velocities (for the parity test)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "bubbles_api.h"

using namespace bbx;

int main(int argc, char **argv){
    Float domainScaling = 0.6f, jitter = 0.001, dt = 0;
    int frames = 2, steps = 0, emit = 0; bool sph = false, mapEmit = false, keyframes = false, meshObstacle = false;
    std::string out, dump, load;
    for(int i = 1; i < argc; i++){
        std::string a = argv[i];
        auto next = [&](){ if(i + 1 >= argc){ std::fprintf(stderr, "missing value for %s\n", a.c_str()); std::exit(2); } return std::string(argv[++i]); };
        if(a == "--scaling") domainScaling = std::atof(next().c_str());
        else if(a == "--frames") frames = std::atoi(next().c_str());
        else if(a == "--steps") steps = std::atoi(next().c_str());
        else if(a == "--dt") dt = std::atof(next().c_str());
        else if(a == "--jitter") jitter = std::atof(next().c_str());
        else if(a == "--out") out = next();
        else if(a == "--dump") dump = next();
        else if(a == "--sph") sph = true;
        else if(a == "--emit") emit = std::atoi(next().c_str());
        else if(a == "--map-emit") mapEmit = true;
        else if(a == "--load") load = next();
        else if(a == "--keyframes") keyframes = true;
        else if(a == "--mesh") meshObstacle = true;
        else{ std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    try{
        std::printf("===== PCISPH Solver 3D -- Dam Break (bbx)\n");
        Float spacing = 0.02f;        // float literals as in the reference scene
        Float spacingScale = 1.8f;
        Float boxFluidLen = 0.5 * domainScaling, boxFluidYLen = 0.9 * domainScaling;
        Float boxLen = 1.3 * domainScaling, boxYLen = 1.2 * domainScaling;
        vec3f containerSize(boxLen, boxYLen, boxLen);
        Float xof = (containerSize.x - boxFluidLen) / 2.0; xof -= spacing;
        Float zof = (containerSize.z - boxFluidLen) / 2.0; zof -= spacing;
        Float yof = (containerSize.y - boxFluidYLen) / 2.0; yof -= spacing;
        vec3f boxSize(boxFluidLen, boxFluidYLen, boxFluidLen);

        ShapePtr container = MakeBox(Transform(), containerSize, true);
        ShapePtr boxp = MakeBox(Translate(xof, -yof, zof), boxSize);

        ContinuousParticleSetBuilder3 pBuilder(2500000);   // (room for --emit; otherwise the same as ParticleSetBuilder3)
        VolumeParticleEmitterSet3 emitterSet;
        VolumeParticleEmitter3 emitterp(boxp, boxp->GetBounds(), spacing, vec3f(0, -6, 0));
        emitterSet.AddEmitter(&emitterp);
        emitterSet.SetJitter(jitter);
        if(load.empty()) emitterSet.Emit(&pBuilder);
        else{
            ParticleSetBuilder3 fromFile; int flags = SERIALIZER_POSITION;
            const int n = SerializerLoadSphDataSet3(&fromFile, load.c_str(), flags);
            if(n <= 0){ std::fprintf(stderr, "no particles in %s\n", load.c_str()); return 1; }
            for(int i = 0; i < n; i++) pBuilder.AddParticle(fromFile.positions[i], (flags & SERIALIZER_VELOCITY) ? fromFile.velocities[i] : vec3f(0, -6, 0));
            pBuilder.Commit();
            std::printf("loaded %d particles (format %s) from %s\n", n, SerializerStringFromFlags(flags).c_str(), load.c_str());
        }

        auto domainGrid = UtilBuildGridForDomain(container->GetBounds(), spacing, spacingScale);
        ColliderSetBuilder3 cBuilder;
        cBuilder.AddCollider3(container);
        // --keyframes: a sphere that slides across the floor and back (collider 1)
        const Float ballR = 0.12 * domainScaling;
        const vec3f ballA(-0.3 * boxLen, -0.5 * boxYLen + ballR, -0.3 * boxLen), ballB(0.3 * boxLen, -0.5 * boxYLen + ballR, 0.1 * boxLen);
        ShapePtr ball = MakeSphere(Translate(ballA), ballR);
        TransformSequence sequence;
        if(keyframes){
            cBuilder.AddCollider3(ball, 0.1);
            Transform k0 = Translate(ballA), k1 = Translate(ballB) * Rotate(180, vec3f(0, 1, 0)), k2 = Translate(vec3f(ballB.x, ballB.y, ballA.z)) * Rotate(270, vec3f(0, 1, 0));
            sequence.AddInterpolation(&k0, &k1, 0, 1);
            sequence.AddInterpolation(&k1, &k2, 1, 2);
            sequence.AddRestore(2, 3);
        }
        if(meshObstacle){
            const Float r = 0.15 * domainScaling; const vec3f c(-0.15 * boxLen, -0.5 * boxYLen + r, 0.15 * boxLen);
            std::vector<vec3f> pts = {c + vec3f(r, 0, 0), c + vec3f(-r, 0, 0), c + vec3f(0, r, 0), c + vec3f(0, -r, 0), c + vec3f(0, 0, r), c + vec3f(0, 0, -r)};
            std::vector<int> tri = {0, 2, 4, 2, 1, 4, 1, 3, 4, 3, 0, 4, 2, 0, 5, 1, 2, 5, 3, 1, 5, 0, 3, 5};
            cBuilder.AddCollider3(MakeMesh(pts, tri), 0.0);
        }
        auto colliders = cBuilder.GetColliderSet();
        colliders->GenerateSDFs();   // (bakes the grid of mesh colliders; nothing to do for the others)

        auto sphSet = SphParticleSet3FromContinuousBuilder(&pBuilder);
        if(!emit && !mapEmit) sphSet->reservedSize = 0;  // no emission: size the engine for the block alone
        sphSet->SetRelativeKernelRadius(spacingScale);
        std::printf("particles %d, cells %d (%d x %d x %d)\n", pBuilder.GetParticleCount(), domainGrid->desc.total,
                    domainGrid->desc.n[0], domainGrid->desc.n[1], domainGrid->desc.n[2]);

        PciSphSolver3 pci; SphSolver3 sphSolver;
        auto run = [&](auto &solver){
            solver.Initialize(DefaultSphSolverData3());
            solver.Setup(WaterDensity, spacing, spacingScale, domainGrid, sphSet);
            solver.SetColliders(colliders);
            if(mapEmit) pBuilder.MapGrid(domainGrid);
            auto save = [&](int frame){
                if(out.empty()) return;
                std::string path = out + "/out_" + std::to_string(frame) + ".txt";
                std::remove(path.c_str());
                SerializerSaveSphDataSet3(solver.GetSphSolverData(), path.c_str(), SERIALIZER_POSITION);
            };
            if(steps > 0){
                if(!(dt > 0)) dt = sph ? 1.44e-4 : 7.2e-4;
                save(0);
                solver.AdvanceTimeStep(dt, steps);
                save(1);
                std::printf("%d fixed sub-steps of %g s\n", steps, dt);
            }else{
                Float targetInterval = 1.0 / 240.0;
                RunSimulation3(&solver, targetInterval, [&](int step) -> int {
                    if(step == 0){ save(0); return 1; }
                    std::printf("Step (%d) : %g ms - Particles %d\n", step - 1, solver.GetAdvanceTime(), solver.GetParticleCount());
                    save(step);
                    if(keyframes){
                        // one loop of the sequence every 120 frames; velocities = displacement / rotation since the last frame over
                        // the frame time (test_pcisph3.cpp:124-134)
                        const int loop = 120;
                        Float f = 3 * ((Float)(step % loop)) / (Float)loop;
                        Transform transform; vec3f linear, angular;
                        sequence.Interpolate(f, &transform, &linear, &angular);
                        ball->Update(transform);
                        ball->SetVelocities(linear * (1.0 / targetInterval), angular * (1.0 / targetInterval));
                        solver.UpdateCollider(1);
                    }
                    if(emit && step < frames){
                        // a sheet of K particles one spacing apart, dropped from above the block
                        int side = 1; while(side * side < emit) side++;
                        for(int q = 0; q < emit; q++)
                            pBuilder.AddParticle(vec3f(xof - 0.5 * boxFluidLen + spacing * (q % side + 1), 0.5 * boxYLen - 3 * spacing, zof - 0.5 * boxFluidLen + spacing * (q / side + 1)), vec3f(0, -3, 0));
                        pBuilder.Commit();
                    }
                    if(mapEmit && step < frames){
                        const int added = pBuilder.MapGridEmit([](const vec3f &) { return vec3f(0, -6, 0); }, spacing);
                        std::printf("map-emit added %d\n", added);
                    }
                    return step >= frames ? 0 : 1;
                });
            }
            bbx_step_stats st = solver.Stats();
            std::printf("substeps %d, neighbour overflow %d, clamped %d, non-finite %d\n", st.substeps, st.neighbor_overflow, st.clamped, st.nan_count);
        };
        if(sph) run(sphSolver); else run(pci);

        ParticleSet3 *ps = sphSet->GetParticleSet();
        vec3f p0 = ps->GetParticlePosition(0);
        std::printf("p[0] = %.17g %.17g %.17g  rho[0] = %.9g\n", p0.x, p0.y, p0.z, ps->GetParticleDensity(0));
        if(!dump.empty()){
            FILE *fp = std::fopen(dump.c_str(), "wb");
            if(!fp){ std::fprintf(stderr, "cannot open %s\n", dump.c_str()); return 1; }
            double n = ps->GetParticleCount();
            std::fwrite(&n, sizeof(double), 1, fp);
            std::fwrite(ps->positions.data(), sizeof(vec3f), ps->positions.size(), fp);
            std::fwrite(ps->velocities.data(), sizeof(vec3f), ps->velocities.size(), fp);
            std::fclose(fp);
        }
        std::printf("===== OK\n");
    }catch(const std::exception &ex){
        std::fprintf(stderr, "%s\n", ex.what());
        return 1;
    }
    return 0;
}
