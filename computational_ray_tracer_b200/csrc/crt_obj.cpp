// crt_obj.cpp -- Wavefront OBJ ingestion into the MeshCache::Model layout (RayTracer/AssetManager.h:20-47).
//
// Stands in for the reference's ASSIMPLoader (RayTracer/AssetManager.cpp:67-190), which reads any assimp format with
// aiProcess_Triangulate | aiProcess_CalcTangentSpace | aiProcess_GenNormals and no vertex joining.  assimp is not part of
// the reference repository, so this is a restatement of what that call produces for OBJ input, not a parity-pinned port:
//   * one mesh per object / group / material run (assimp: one aiMesh per material of each object),
//   * one vertex per face corner (no JoinIdenticalVertices), indices 0..3T-1,
//   * polygons fan-triangulated (v0, vi, vi+1),
//   * normals from `vn` when the face references them, flat face normals otherwise (GenNormals).
// Only positions, normals and indices are kept (what the render path reads).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "crt_host.h"

using namespace crt;

struct crt_obj {
    struct Mesh { std::string name; std::vector<float> pos, nrm; std::vector<uint32_t> idx; };
    std::vector<Mesh> meshes;
};

namespace {
struct Corner { int v, vn; };
// "v", "v/vt", "v//vn", "v/vt/vn"; negative indices are relative to the end (OBJ spec)
bool parse_corner(const char*& p, int nv, int nn, Corner& c) {
    char* end;
    long v = std::strtol(p, &end, 10);
    if (end == p) return false;
    p = end;
    long vn = 0;
    if (*p == '/') {
        ++p;
        if (*p != '/') { std::strtol(p, &end, 10); p = end; }          // vt ignored
        if (*p == '/') { ++p; vn = std::strtol(p, &end, 10); p = end; }
    }
    c.v = (int)(v < 0 ? nv + v : v - 1);
    c.vn = vn == 0 ? -1 : (int)(vn < 0 ? nn + vn : vn - 1);
    return c.v >= 0 && c.v < nv && c.vn < nn;
}
}  // namespace

extern "C" {

int crt_obj_load(const char* path, crt_obj** out) {
    if (!path || !out) { set_error("obj_load: bad arguments"); return 1; }
    std::FILE* f = std::fopen(path, "r");
    if (!f) { set_error(std::string("obj_load: cannot open ") + path); return 1; }
    auto* o = new crt_obj;
    std::vector<float> V, N;
    std::string pending = "default";
    bool need_new = true;
    char line[4096];
    long lineno = 0;
    while (std::fgets(line, sizeof line, f)) {
        ++lineno;
        const char* p = line;
        while (*p == ' ' || *p == '\t') ++p;
        if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
            float x, y, z;
            if (std::sscanf(p + 1, "%f %f %f", &x, &y, &z) != 3) { set_error("obj_load: bad vertex at line " + std::to_string(lineno)); std::fclose(f); delete o; return 1; }
            V.push_back(x); V.push_back(y); V.push_back(z);
        } else if (p[0] == 'v' && p[1] == 'n') {
            float x, y, z;
            if (std::sscanf(p + 2, "%f %f %f", &x, &y, &z) != 3) { set_error("obj_load: bad normal at line " + std::to_string(lineno)); std::fclose(f); delete o; return 1; }
            N.push_back(x); N.push_back(y); N.push_back(z);
        } else if ((p[0] == 'o' || p[0] == 'g') && (p[1] == ' ' || p[1] == '\t')) {
            pending = std::string(p + 2);
            while (!pending.empty() && (pending.back() == '\n' || pending.back() == '\r' || pending.back() == ' ')) pending.pop_back();
            need_new = true;
        } else if (std::strncmp(p, "usemtl", 6) == 0) {
            need_new = true;
        } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
            std::vector<Corner> cs;
            const char* q = p + 1;
            while (true) {
                while (*q == ' ' || *q == '\t') ++q;
                if (*q == '\0' || *q == '\n' || *q == '\r' || *q == '#') break;
                Corner c;
                if (!parse_corner(q, (int)(V.size() / 3), (int)(N.size() / 3), c)) { set_error("obj_load: bad face at line " + std::to_string(lineno)); std::fclose(f); delete o; return 1; }
                cs.push_back(c);
            }
            if (cs.size() < 3) continue;
            if (need_new || o->meshes.empty()) {
                if (o->meshes.empty() || !o->meshes.back().idx.empty()) o->meshes.emplace_back();
                o->meshes.back().name = pending;
                need_new = false;
            }
            crt_obj::Mesh& m = o->meshes.back();
            for (size_t k = 1; k + 1 < cs.size(); ++k) {
                const Corner tri[3] = {cs[0], cs[k], cs[k + 1]};
                f3 P[3];
                for (int a = 0; a < 3; ++a) P[a] = mk3(V[3 * tri[a].v], V[3 * tri[a].v + 1], V[3 * tri[a].v + 2]);
                f3 fn = cross3(P[1] - P[0], P[2] - P[0]);
                float len = length3(fn);
                fn = len > 0 ? fn * (1.0f / len) : mk3(0, 0, 0);
                for (int a = 0; a < 3; ++a) {
                    m.idx.push_back((uint32_t)(m.pos.size() / 3));
                    m.pos.push_back(P[a].x); m.pos.push_back(P[a].y); m.pos.push_back(P[a].z);
                    if (tri[a].vn >= 0) { m.nrm.push_back(N[3 * tri[a].vn]); m.nrm.push_back(N[3 * tri[a].vn + 1]); m.nrm.push_back(N[3 * tri[a].vn + 2]); }
                    else { m.nrm.push_back(fn.x); m.nrm.push_back(fn.y); m.nrm.push_back(fn.z); }
                }
            }
        }
    }
    std::fclose(f);
    if (!o->meshes.empty() && o->meshes.back().idx.empty()) o->meshes.pop_back();
    if (o->meshes.empty()) { set_error(std::string("obj_load: no faces in ") + path); delete o; return 1; }
    *out = o;
    return 0;
}
void crt_obj_destroy(crt_obj* o) { delete o; }
int crt_obj_mesh_count(const crt_obj* o) { return o ? (int)o->meshes.size() : 0; }
int crt_obj_mesh_info(const crt_obj* o, int i, uint32_t* n_vertices, uint32_t* n_triangles, char* name, int name_cap) {
    if (!o || i < 0 || i >= (int)o->meshes.size()) { set_error("obj_mesh_info: bad mesh index"); return 1; }
    const crt_obj::Mesh& m = o->meshes[i];
    *n_vertices = (uint32_t)(m.pos.size() / 3); *n_triangles = (uint32_t)(m.idx.size() / 3);
    if (name && name_cap > 0) { std::strncpy(name, m.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
    return 0;
}
int crt_obj_mesh_copy(const crt_obj* o, int i, float* positions, float* normals, uint32_t* indices) {
    if (!o || i < 0 || i >= (int)o->meshes.size()) { set_error("obj_mesh_copy: bad mesh index"); return 1; }
    const crt_obj::Mesh& m = o->meshes[i];
    if (positions) std::memcpy(positions, m.pos.data(), m.pos.size() * sizeof(float));
    if (normals) std::memcpy(normals, m.nrm.data(), m.nrm.size() * sizeof(float));
    if (indices) std::memcpy(indices, m.idx.data(), m.idx.size() * sizeof(uint32_t));
    return 0;
}

}  // extern "C"
