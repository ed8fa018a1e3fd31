// frame_tool: writes one bbtool-readable text frame with the facade's SerializerSaveSphDataSet3 from a raw state
// file -- host only (no engine), used by tests/test_frame_writer.py to compare the writer byte for byte with the
// reference's own (src/third/serializer.cpp:812-921) and to feed the reference's reader.
//   frame_tool <state.bin> <out.txt> <flags> [--box tx ty tz sx sy sz | --sphere tx ty tz r] ...
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

int main(int argc, char **argv){
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
    bbx::ColliderSetBuilder3 cb; int shapes = 0;
    for(int a = 4; a < argc; ){
        std::string k = argv[a];
        if(k == "--box" && a + 6 < argc){
            cb.AddCollider3(bbx::MakeBox(bbx::Translate(std::atof(argv[a+1]), std::atof(argv[a+2]), std::atof(argv[a+3])),
                                         bbx::vec3f(std::atof(argv[a+4]), std::atof(argv[a+5]), std::atof(argv[a+6])))); a += 7; shapes++;
        }else if(k == "--sphere" && a + 4 < argc){
            cb.AddCollider3(bbx::MakeSphere(bbx::Translate(std::atof(argv[a+1]), std::atof(argv[a+2]), std::atof(argv[a+3])), std::atof(argv[a+4]))); a += 5; shapes++;
        }else{ std::fprintf(stderr, "bad shape argument %s\n", argv[a]); return 2; }
    }
    if(shapes) bbx::UtilSaveSimulation3(cb.GetColliderSet().get(), data.get(), argv[2], std::atoi(argv[3]));
    else bbx::SerializerSaveSphDataSet3(data.get(), argv[2], std::atoi(argv[3]));
    return 0;
}
