// frame_tool: writes one bbtool-readable text frame with the facade's SerializerSaveSphDataSet3 from a raw state
// file -- host only (no engine), used by tests/test_frame_writer.py to compare the writer byte for byte with the
// reference's own (src/third/serializer.cpp:812-921) and to feed the reference's reader.
//   frame_tool <state.bin> <out.txt> <flags> [--boundary b.bin n.bin] [--box tx ty tz sx sy sz | --sphere tx ty tz r] ...
//     --boundary: per-particle boundary layer (int64 n, double v0[n]) and normals (int64 n, double normal[3n]); without
//     shapes the frame is written with the boundary vector UtilGetBoundaryState builds from them
//     with shapes: UtilSaveSimulation3 (shape blocks of every collider but the last + the particle block)
//   state.bin: int64 n, double spacing, double mass, double pos[3n], double vel[3n], double rho[n]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>
#include "bubbles_api.h"

// frame_tool --load <frame.txt> <out.bin>: the facade's reader (SerializerLoadParticles3) on a frame file;
//   out.bin: int64 n, int64 flags, double pos[3n], vel[3n], rho[n], mass[n]
static int load_mode(const char *in, const char *out){
    std::vector<bbx::SerializedParticle> ps; int flags = 0;
    const int n = bbx::SerializerLoadParticles3(&ps, in, flags);
    if(n < 0) return 1;
    FILE *fp = std::fopen(out, "wb");
    if(!fp) return 1;
    int64_t hdr[2] = {n, flags};
    std::fwrite(hdr, sizeof(int64_t), 2, fp);
    for(int pass = 0; pass < 4; pass++) for(int i = 0; i < n; i++){
        const bbx::SerializedParticle &q = ps[(size_t)i];
        if(pass == 0){ double v[3] = {q.position.x, q.position.y, q.position.z}; std::fwrite(v, 8, 3, fp); }
        else if(pass == 1){ double v[3] = {q.velocity.x, q.velocity.y, q.velocity.z}; std::fwrite(v, 8, 3, fp); }
        else if(pass == 2){ double v = q.density; std::fwrite(&v, 8, 1, fp); }
        else{ double v = (flags & bbx::SERIALIZER_MASS) ? q.mass : 0.0; std::fwrite(&v, 8, 1, fp); }
    }
    std::fclose(fp);
    return 0;
}

// frame_tool --emit <out.bin> tx ty tz sx sy sz spacing vx vy vz jitter seed: the facade's VolumeParticleEmitter3 over a
//   box (libc rand() seeded like the reference scene scripts do); out.bin: int64 n, double pos[3n], vel[3n]
static int emit_mode(int argc, char **argv){
    if(argc != 15) return 2;
    double a[12]; for(int k = 0; k < 12; k++) a[k] = std::atof(argv[3 + k]);
    std::srand((unsigned)a[11]);
    auto box = bbx::MakeBox(bbx::Translate(a[0], a[1], a[2]), bbx::vec3f(a[3], a[4], a[5]));
    bbx::VolumeParticleEmitter3 em(box, box->GetBounds(), a[6], bbx::vec3f(a[7], a[8], a[9]));
    bbx::VolumeParticleEmitterSet3 set; set.AddEmitter(&em); set.SetJitter(a[10]);
    bbx::ParticleSetBuilder3 b; set.Emit(&b);
    FILE *fp = std::fopen(argv[2], "wb");
    if(!fp) return 1;
    int64_t n = b.GetParticleCount();
    std::fwrite(&n, sizeof(n), 1, fp);
    std::fwrite(b.positions.data(), sizeof(bbx::vec3f), b.positions.size(), fp);
    std::fwrite(b.velocities.data(), sizeof(bbx::vec3f), b.velocities.size(), fp);
    std::fclose(fp);
    return 0;
}

// frame_tool --mapemit <in.bin> <out.bin>: the facade's MapGridEmit rule (MapGridEmitCandidates) as a pure function.
//   in.bin: double d; int64 n, total, m; double pos[3n]; int32 cell_count[total]; int32 cell_order[n];
//           then m mapped cells: int64 cell, int64 k, double tmpl[3k]
//   out.bin: int64 count, double new_pos[3 count]
static int mapemit_mode(const char *in, const char *out){
    FILE *fp = std::fopen(in, "rb");
    if(!fp) return 1;
    double d = 0; int64_t n = 0, total = 0, m = 0;
    size_t r = std::fread(&d, 8, 1, fp) + std::fread(&n, 8, 1, fp) + std::fread(&total, 8, 1, fp) + std::fread(&m, 8, 1, fp);
    std::vector<bbx::vec3f> pos((size_t)n); std::vector<int> count((size_t)total), order((size_t)n);
    r += std::fread(pos.data(), sizeof(bbx::vec3f), pos.size(), fp) + std::fread(count.data(), 4, count.size(), fp) + std::fread(order.data(), 4, order.size(), fp);
    std::map<unsigned, std::vector<bbx::vec3f>> mapped;
    for(int64_t c = 0; c < m; c++){
        int64_t cell = 0, k = 0; r += std::fread(&cell, 8, 1, fp) + std::fread(&k, 8, 1, fp);
        std::vector<bbx::vec3f> &v = mapped[(unsigned)cell]; v.resize((size_t)k);
        r += std::fread(v.data(), sizeof(bbx::vec3f), v.size(), fp);
    }
    std::fclose(fp);
    (void)r;
    std::vector<bbx::vec3f> add = bbx::MapGridEmitCandidates(mapped, (int)total, count.data(), order.data(), pos.data(), d);
    fp = std::fopen(out, "wb");
    if(!fp) return 1;
    int64_t cnt = (int64_t)add.size();
    std::fwrite(&cnt, 8, 1, fp); std::fwrite(add.data(), sizeof(bbx::vec3f), add.size(), fp);
    std::fclose(fp);
    return 0;
}

// frame_tool --tseq <job.txt> <out.bin>: the facade's TransformSequence / QuaternionSequence driven by the same job lines
//   the reference harness takes (tseq_add, tseq_restore, tseq_eval, qseq_add, qseq_eval; the eval prefix argument is ignored).
//   out.bin, per eval command in order: int64 n, double m[16 n], minv[16 n], linear[3 n], angular[3 n]
static int tseq_mode(const char *job, const char *out){
    std::ifstream in(job);
    if(!in) return 1;
    FILE *fp = std::fopen(out, "wb");
    if(!fp) return 1;
    bbx::TransformSequence tseq; bbx::QuaternionSequence qseq;
    std::string cmd;
    while(in >> cmd){
        if(cmd == "tseq_add"){
            double a[16]; for(int k = 0; k < 16; k++) in >> a[k];
            double s0, s1; in >> s0 >> s1;
            bbx::Transform k0 = bbx::Translate(bbx::vec3f(a[0], a[1], a[2])) * bbx::Rotate(a[3], bbx::vec3f(a[4], a[5], a[6])) * bbx::Scale(a[7]);
            bbx::Transform k1 = bbx::Translate(bbx::vec3f(a[8], a[9], a[10])) * bbx::Rotate(a[11], bbx::vec3f(a[12], a[13], a[14])) * bbx::Scale(a[15]);
            tseq.AddInterpolation(&k0, &k1, s0, s1);
        }else if(cmd == "tseq_restore"){ double s0, s1; in >> s0 >> s1; tseq.AddRestore(s0, s1); }
        else if(cmd == "qseq_add"){ double ang, x, y, z, t; in >> ang >> x >> y >> z >> t; qseq.AddQuaternion(ang, bbx::vec3f(x, y, z), t); }
        else if(cmd == "tseq_eval" || cmd == "qseq_eval"){
            double t0, dt; int64_t n; std::string prefix; in >> t0 >> dt >> n >> prefix;
            std::vector<double> m(16 * (size_t)n), mi(16 * (size_t)n), lin(3 * (size_t)n), ang(3 * (size_t)n);
            for(int64_t i = 0; i < n; i++){
                bbx::Transform tr; bbx::vec3f l(0.0), w(0.0);
                if(cmd == "tseq_eval") tseq.Interpolate(t0 + i * dt, &tr, &l, &w);
                else qseq.Interpolate(t0 + i * dt, &tr, &w);
                for(int r = 0; r < 4; r++) for(int c = 0; c < 4; c++){ m[16 * (size_t)i + 4 * r + c] = tr.m[r][c]; mi[16 * (size_t)i + 4 * r + c] = tr.mInv[r][c]; }
                for(int k = 0; k < 3; k++){ lin[3 * (size_t)i + k] = l[k]; ang[3 * (size_t)i + k] = w[k]; }
            }
            std::fwrite(&n, 8, 1, fp);
            std::fwrite(m.data(), 8, m.size(), fp); std::fwrite(mi.data(), 8, mi.size(), fp);
            std::fwrite(lin.data(), 8, lin.size(), fp); std::fwrite(ang.data(), 8, ang.size(), fp);
        }else{ std::fprintf(stderr, "unknown command %s\n", cmd.c_str()); std::fclose(fp); return 2; }
    }
    std::fclose(fp);
    return 0;
}

// frame_tool --bake <mesh.bin> <out.bin> dx margin [n_queries q.bin]: the facade's MakeMesh + GenerateShapeSDF.
//   mesh.bin: int64 nv, nt; double points[3 nv]; int32 triangles[3 nt]
//   out.bin: int64 res[3]; double spacing, origin[3], bounds[6]; double field[res0 res1 res2]
//   optional q.bin (int64 n, double p[3n]) -> appended: double distance[n] (MeshClosestDistance)
static int bake_mode(int argc, char **argv){
    if(argc != 6 && argc != 7) return 2;
    FILE *fp = std::fopen(argv[2], "rb");
    if(!fp) return 1;
    int64_t nv = 0, nt = 0; size_t r = std::fread(&nv, 8, 1, fp) + std::fread(&nt, 8, 1, fp);
    std::vector<bbx::vec3f> pts((size_t)nv); std::vector<int> tri(3 * (size_t)nt);
    r += std::fread(pts.data(), sizeof(bbx::vec3f), pts.size(), fp) + std::fread(tri.data(), 4, tri.size(), fp);
    std::fclose(fp); (void)r;
    bbx::ShapePtr s = bbx::MakeMesh(pts, tri);
    bbx::GenerateShapeSDF(s.get(), std::atof(argv[4]), std::atof(argv[5]));
    fp = std::fopen(argv[3], "wb");
    if(!fp) return 1;
    int64_t res[3] = {s->sdfResolution[0], s->sdfResolution[1], s->sdfResolution[2]};
    bbx::Bounds3f b = s->GetBounds();
    double meta[10] = {s->sdfSpacing, s->sdfOrigin.x, s->sdfOrigin.y, s->sdfOrigin.z, b.pMin.x, b.pMin.y, b.pMin.z, b.pMax.x, b.pMax.y, b.pMax.z};
    std::fwrite(res, 8, 3, fp); std::fwrite(meta, 8, 10, fp);
    std::fwrite(s->sdfField.data(), 8, s->sdfField.size(), fp);
    if(argc == 7){
        FILE *fq = std::fopen(argv[6], "rb");
        if(!fq){ std::fclose(fp); return 1; }
        int64_t n = 0; r = std::fread(&n, 8, 1, fq);
        std::vector<bbx::vec3f> q((size_t)n); r += std::fread(q.data(), sizeof(bbx::vec3f), q.size(), fq); std::fclose(fq);
        for(int64_t i = 0; i < n; i++){ double d = bbx::MeshClosestDistance(*s, q[(size_t)i]); std::fwrite(&d, 8, 1, fp); }
    }
    std::fclose(fp);
    return 0;
}

// frame_tool --obj <file.obj> <out.bin>: the facade's LoadObj; out.bin: int64 nv, nt; double points[3 nv]; int32 triangles[3 nt]
static int obj_mode(const char *in, const char *out){
    bbx::ParsedMesh m = bbx::LoadObj(in);
    FILE *fp = std::fopen(out, "wb");
    if(!fp) return 1;
    int64_t hdr[2] = {m.nVertices, m.nTriangles};
    std::fwrite(hdr, 8, 2, fp);
    std::fwrite(m.p.data(), sizeof(bbx::vec3f), m.p.size(), fp);
    std::fwrite(m.indices.data(), 4, m.indices.size(), fp);
    std::fclose(fp);
    return 0;
}

int main(int argc, char **argv){
    if(argc == 4 && std::string(argv[1]) == "--obj") return obj_mode(argv[2], argv[3]);
    if(argc >= 2 && std::string(argv[1]) == "--bake") return bake_mode(argc, argv);
    if(argc == 4 && std::string(argv[1]) == "--tseq") return tseq_mode(argv[2], argv[3]);
    if(argc == 4 && std::string(argv[1]) == "--load") return load_mode(argv[2], argv[3]);
    if(argc == 4 && std::string(argv[1]) == "--mapemit") return mapemit_mode(argv[2], argv[3]);
    if(argc >= 2 && std::string(argv[1]) == "--emit") return emit_mode(argc, argv);
    if(argc < 4){ std::fprintf(stderr, "usage: frame_tool state.bin out.txt flags\n"); return 2; }
    FILE *fp = std::fopen(argv[1], "rb");
    if(!fp){ std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
    int64_t n = 0; double spacing = 0, mass = 0;
    size_t r = std::fread(&n, sizeof(n), 1, fp) + std::fread(&spacing, 8, 1, fp) + std::fread(&mass, 8, 1, fp);
    std::vector<double> pos(3 * (size_t)n), vel(3 * (size_t)n), rho((size_t)n);
    r += std::fread(pos.data(), 8, pos.size(), fp) + std::fread(vel.data(), 8, vel.size(), fp) + std::fread(rho.data(), 8, rho.size(), fp);
    std::fclose(fp);
    if(r != 3 + 7 * (size_t)n){ std::fprintf(stderr, "short state file\n"); return 2; }
    bbx::ParticleSetBuilder3 b;
    for(int64_t i = 0; i < n; i++) b.AddParticle(bbx::vec3f(pos[3*i], pos[3*i+1], pos[3*i+2]), bbx::vec3f(vel[3*i], vel[3*i+1], vel[3*i+2]));
    auto set = bbx::SphParticleSet3FromBuilder(&b);
    set->SetTargetSpacing(spacing);
    set->set.mass = mass;
    for(int64_t i = 0; i < n; i++) set->set.densities[(size_t)i] = rho[(size_t)i];
    auto data = bbx::DefaultSphSolverData3(true);
    data->sphpSet = set;
    bbx::ColliderSetBuilder3 cb; int shapes = 0; bool withBoundary = false;
    for(int a = 4; a < argc; ){
        std::string k = argv[a];
        if(k == "--boundary" && a + 2 < argc){
            FILE *fb = std::fopen(argv[a + 1], "rb"), *fn = std::fopen(argv[a + 2], "rb");
            if(!fb || !fn){ std::fprintf(stderr, "cannot open the boundary files\n"); return 2; }
            int64_t nb = 0; size_t q = std::fread(&nb, 8, 1, fb);
            std::vector<double> v0((size_t)nb); q += std::fread(v0.data(), 8, v0.size(), fb); std::fclose(fb);
            q += std::fread(&nb, 8, 1, fn);
            std::vector<double> nr(3 * (size_t)nb); q += std::fread(nr.data(), 8, nr.size(), fn); std::fclose(fn); (void)q;
            for(int64_t i = 0; i < nb; i++){ set->set.SetParticleV0((int)i, v0[(size_t)i]); set->set.SetParticleNormal((int)i, bbx::vec3f(nr[3*i], nr[3*i+1], nr[3*i+2])); }
            withBoundary = true; a += 3;
        }else if(k == "--box" && a + 6 < argc){
            cb.AddCollider3(bbx::MakeBox(bbx::Translate(std::atof(argv[a+1]), std::atof(argv[a+2]), std::atof(argv[a+3])),
                                         bbx::vec3f(std::atof(argv[a+4]), std::atof(argv[a+5]), std::atof(argv[a+6])))); a += 7; shapes++;
        }else if(k == "--sphere" && a + 4 < argc){
            cb.AddCollider3(bbx::MakeSphere(bbx::Translate(std::atof(argv[a+1]), std::atof(argv[a+2]), std::atof(argv[a+3])), std::atof(argv[a+4]))); a += 5; shapes++;
        }else{ std::fprintf(stderr, "bad shape argument %s\n", argv[a]); return 2; }
    }
    if(shapes) bbx::UtilSaveSimulation3(cb.GetColliderSet().get(), data.get(), argv[2], std::atoi(argv[3]));
    else if(withBoundary){
        std::vector<int> boundaries; bbx::UtilGetBoundaryState(&set->set, &boundaries);
        bbx::SerializerSaveSphDataSet3(data.get(), argv[2], std::atoi(argv[3]), &boundaries);
    }
    else bbx::SerializerSaveSphDataSet3(data.get(), argv[2], std::atoi(argv[3]));
    return 0;
}
