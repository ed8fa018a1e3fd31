// Wavefront .obj -> triangle mesh, with the reference's loader semantics (included at the end of bubbles_api.h, namespace bbx):
//
//   reference (src/third/obj_loader.cpp)                  here
//   LoadObj(path) :434-441, :443-625                      LoadObj(path): `v` lines through ParseV3 / ParseFloat (the same digit
//                                                          loop as the frame reader -> the same bits), `f` lines with 3 or 4
//                                                          corners (a quad becomes 0 1 2 + 0 2 3), corners as i, i/j, i//k,
//                                                          i/j/k, negative indices relative to the vertices read so far
//   parseTriple / fixIndex :172-246                       ParseCorner
//   FillMesh :335-421                                     vertices re-numbered in the order the faces first use them (unused
//                                                          vertices are dropped), exactly as mesh->p / mesh->indices[].x come out
//   MakeMesh(path | mesh, toWorld) shapes/bvh.cpp:51-61   MakeMesh(const ParsedMesh &, toWorld): Transform::Mesh moves the points
//
// Normals, texture coordinates, materials and `o` / `g` splitting are not kept: a collider needs the surface only.
// tests/test_mesh_sdf.py: points and triangle indices equal the reference loader's on a file that uses every corner syntax.
#pragma once

struct ParsedMesh {
    std::vector<vec3f> p;        // vertices in first-use order
    std::vector<int> indices;    // 3 vertex indices per triangle
    int nVertices = 0, nTriangles = 0;
    std::string name;
};

// one face corner: returns false on a malformed corner (index 0, or nothing to read)
inline bool ParseCorner(const char **token, int vsize, int *v_idx){
    auto fix = [](int idx, int n, int *ret){ if(idx > 0){ *ret = idx - 1; return true; } if(idx < 0){ *ret = n + idx; return true; } return false; };
    if(!fix(std::atoi(*token), vsize, v_idx)) return false;
    *token += std::strcspn(*token, "/ \t\r");
    if((*token)[0] != '/') return true;
    (*token)++;
    if((*token)[0] == '/'){ (*token)++; *token += std::strcspn(*token, "/ \t\r"); return true; }        // i//k
    *token += std::strcspn(*token, "/ \t\r");                                                            // i/j
    if((*token)[0] != '/') return true;
    (*token)++; *token += std::strcspn(*token, "/ \t\r");                                                // i/j/k
    return true;
}

inline ParsedMesh LoadObj(const char *path){
    ParsedMesh mesh;
    { const char *slash = std::strrchr(path, '/'); mesh.name = slash ? slash + 1 : path; }
    std::ifstream ifs(path);
    if(!ifs){ std::printf("[OBJ LOADER] Could not open file %s\n", path); return mesh; }
    std::vector<vec3f> v; std::vector<int> corners;
    std::string line;
    while(std::getline(ifs, line)){
        if(!line.empty() && line.back() == '\r') line.pop_back();
        if(line.empty()) continue;
        const char *token = line.c_str();
        token += std::strspn(token, " \t");
        if(token[0] == '\0' || token[0] == '#') continue;
        auto space = [](char c){ return c == ' ' || c == '\t'; };
        if(token[0] == 'v' && space(token[1])){ token += 2; v.push_back(ParseV3(&token)); continue; }
        if(token[0] == 'f' && space(token[1])){
            token += 2; token += std::strspn(token, " \t");
            int face[4], facen = 0;
            while(!(token[0] == '\r' || token[0] == '\n' || token[0] == '\0')){
                int vi = -1;
                if(!ParseCorner(&token, (int)v.size(), &vi)){ std::printf("[OBJ LOADER] Failed parsing face\n"); break; }
                token += std::strspn(token, " \t\r");
                if(facen >= 4) throw std::runtime_error("[OBJ LOADER] Error: Not a supported face description");   // exit(0) in the reference
                face[facen++] = vi;
            }
            if(facen == 3){ corners.insert(corners.end(), {face[0], face[1], face[2]}); }
            else if(facen == 4){ corners.insert(corners.end(), {face[0], face[1], face[2], face[0], face[2], face[3]}); }
            else std::printf("[OBJ LOADER] Warning unsupported face with %d vertices\n", facen);
        }
    }
    // FillMesh: vertices in the order the corners first use them
    std::vector<int> picked(v.size(), -1);
    for(int c : corners){
        if(c < 0 || (size_t)c >= v.size()) throw std::runtime_error("[OBJ LOADER] face refers to a vertex that does not exist");
        if(picked[(size_t)c] == -1){ picked[(size_t)c] = (int)mesh.p.size(); mesh.p.push_back(v[(size_t)c]); }
        mesh.indices.push_back(picked[(size_t)c]);
    }
    mesh.nVertices = (int)mesh.p.size(); mesh.nTriangles = (int)(mesh.indices.size() / 3);
    return mesh;
}

// MakeMesh(mesh, toWorld) / MakeMesh(path, toWorld): the points move to world space (Transform::Mesh, transform.cpp:405-417)
inline ShapePtr MakeMesh(const ParsedMesh &mesh, const Transform &toWorld, bool reverseOrientation = false){
    std::vector<vec3f> pts(mesh.p.size());
    for(size_t i = 0; i < pts.size(); i++) pts[i] = toWorld.Point(mesh.p[i]);
    return MakeMesh(pts, mesh.indices, reverseOrientation);
}
inline ShapePtr MakeMesh(const char *path, const Transform &toWorld, bool reverseOrientation = false){
    return MakeMesh(LoadObj(path), toWorld, reverseOrientation);
}
