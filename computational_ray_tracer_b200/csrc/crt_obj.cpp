// crt_obj.cpp -- Wavefront OBJ ingestion into the MeshCache::Model layout (RayTracer/AssetManager.h:20-47).
//
// Stands in for the reference's ASSIMPLoader (RayTracer/AssetManager.cpp:67-190), which reads any assimp format with
// aiProcess_Triangulate | aiProcess_CalcTangentSpace | aiProcess_GenNormals and no vertex joining.  assimp is not part of
// the reference repository, so this is a restatement of what that call produces for OBJ input, not a parity-pinned port:
//   * one mesh per object / group / material run (assimp: one aiMesh per material of each object),
//   * one vertex per face corner (no JoinIdenticalVertices), indices 0..3T-1,
//   * polygons fan-triangulated (v0, vi, vi+1),
//   * normals from `vn` when the face references them, flat face normals otherwise (GenNormals),
//   * texture coordinates from `vt` (a mesh has them when every one of its corners references one: aiMesh::mTextureCoords[0]),
//   * tangents / bitangents for meshes with texture coordinates, by assimp's published CalcTangentSpace rule (CalcTangentsProcess:
//     per face T = (w * sy - v * ty) * dir, B = (w * sx - v * tx) * dir with v = p1 - p0, w = p2 - p0, (sx, sy) = uv1 - uv0,
//     (tx, ty) = uv2 - uv0, dir = sign(tx * sy - ty * sx); per corner the part orthogonal to the vertex normal, normalised).  No
//     cross-vertex smoothing: without JoinIdenticalVertices every corner is its own vertex.
// ASSIMPLoader::Process_Mesh then stores the TANGENT a second time where the bitangent belongs (AssetManager.cpp:150-153,
// `modeldata.bitangents.push_back(tangents)`): crt_obj_mesh_attributes reproduces that, so MeshCache::Mesh::bitangents == tangents.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "crt_host.h"

using namespace crt;

struct crt_obj {
    struct Mesh { std::string name; std::vector<float> pos, nrm, uv, tan; std::vector<uint32_t> idx; bool has_uv = true; };
    std::vector<Mesh> meshes;
};

namespace {
struct Corner { int v, vt, vn; };
// "v", "v/vt", "v//vn", "v/vt/vn"; negative indices are relative to the end (OBJ spec)
bool parse_corner(const char*& p, int nv, int nt, int nn, Corner& c) {
    char* end;
    long v = std::strtol(p, &end, 10);
    if (end == p) return false;
    p = end;
    long vt = 0, vn = 0;
    if (*p == '/') {
        ++p;
        if (*p != '/') { vt = std::strtol(p, &end, 10); p = end; }
        if (*p == '/') { ++p; vn = std::strtol(p, &end, 10); p = end; }
    }
    c.v = (int)(v < 0 ? nv + v : v - 1);
    c.vt = vt == 0 ? -1 : (int)(vt < 0 ? nt + vt : vt - 1);
    c.vn = vn == 0 ? -1 : (int)(vn < 0 ? nn + vn : vn - 1);
    if (c.vt < 0 || c.vt >= nt) c.vt = -1;          // a texture index the file never defined: the corner simply has no uv
    return c.v >= 0 && c.v < nv && c.vn < nn;
}
// the part of t orthogonal to n, normalised (assimp: localTangent = tangent - normal * (tangent . normal); NormalizeSafe)
f3 orthonormal_part(f3 t, f3 n) {
    f3 r = t - n * dot3(t, n);
    const float len = length3(r);
    return len > 0 ? r * (1.0f / len) : mk3(0, 0, 0);
}
}  // namespace

extern "C" {

int crt_obj_load(const char* path, crt_obj** out) {
    if (!path || !out) { set_error("obj_load: bad arguments"); return 1; }
    std::FILE* f = std::fopen(path, "r");
    if (!f) { set_error(std::string("obj_load: cannot open ") + path); return 1; }
    auto* o = new crt_obj;
    std::vector<float> V, N, T;
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
        } else if (p[0] == 'v' && p[1] == 't') {
            float u = 0, v = 0;
            if (std::sscanf(p + 2, "%f %f", &u, &v) < 1) { set_error("obj_load: bad texture coordinate at line " + std::to_string(lineno)); std::fclose(f); delete o; return 1; }
            T.push_back(u); T.push_back(v);
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
                if (!parse_corner(q, (int)(V.size() / 3), (int)(T.size() / 2), (int)(N.size() / 3), c)) { set_error("obj_load: bad face at line " + std::to_string(lineno)); std::fclose(f); delete o; return 1; }
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
                // CalcTangentSpace: the face's tangent from its position and uv deltas
                float uvc[3][2] = {{0, 0}, {0, 0}, {0, 0}};
                for (int a = 0; a < 3; ++a) {
                    if (tri[a].vt < 0) m.has_uv = false;
                    else { uvc[a][0] = T[2 * tri[a].vt]; uvc[a][1] = T[2 * tri[a].vt + 1]; }
                }
                const f3 ev = P[1] - P[0], ew = P[2] - P[0];
                float sx = uvc[1][0] - uvc[0][0], sy = uvc[1][1] - uvc[0][1], tx = uvc[2][0] - uvc[0][0], ty = uvc[2][1] - uvc[0][1];
                const float dir = (tx * sy - ty * sx) < 0.0f ? -1.0f : 1.0f;
                if (sx * ty == sy * tx) { sx = 0; sy = 1; tx = 1; ty = 0; }            // degenerate uv: assimp's default directions
                const f3 face_t = (ew * sy - ev * ty) * dir;
                for (int a = 0; a < 3; ++a) {
                    m.idx.push_back((uint32_t)(m.pos.size() / 3));
                    m.pos.push_back(P[a].x); m.pos.push_back(P[a].y); m.pos.push_back(P[a].z);
                    f3 vn = fn;
                    if (tri[a].vn >= 0) vn = mk3(N[3 * tri[a].vn], N[3 * tri[a].vn + 1], N[3 * tri[a].vn + 2]);
                    m.nrm.push_back(vn.x); m.nrm.push_back(vn.y); m.nrm.push_back(vn.z);
                    m.uv.push_back(uvc[a][0]); m.uv.push_back(uvc[a][1]);
                    const f3 lt = orthonormal_part(face_t, vn);
                    m.tan.push_back(lt.x); m.tan.push_back(lt.y); m.tan.push_back(lt.z);
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

// MeshCache::Mesh::texcoords / tangents / bitangents as ASSIMPLoader::Process_Mesh fills them (AssetManager.cpp:104-190).  *available = 1 when
// the mesh has texture coordinates (and therefore tangents); otherwise the arrays are zero-filled like the reference's (:139,:156-157).
int crt_obj_mesh_attributes(const crt_obj* o, int i, float* texcoords, float* tangents, float* bitangents, int* available) {
    if (!o || i < 0 || i >= (int)o->meshes.size()) { set_error("obj_mesh_attributes: bad mesh index"); return 1; }
    const crt_obj::Mesh& m = o->meshes[i];
    if (available) *available = m.has_uv ? 1 : 0;
    const size_t nv = m.pos.size() / 3;
    if (texcoords) { if (m.has_uv) std::memcpy(texcoords, m.uv.data(), 2 * nv * sizeof(float)); else std::memset(texcoords, 0, 2 * nv * sizeof(float)); }
    if (tangents) { if (m.has_uv) std::memcpy(tangents, m.tan.data(), 3 * nv * sizeof(float)); else std::memset(tangents, 0, 3 * nv * sizeof(float)); }
    // the reference pushes `tangents` into the bitangent array (AssetManager.cpp:153)
    if (bitangents) { if (m.has_uv) std::memcpy(bitangents, m.tan.data(), 3 * nv * sizeof(float)); else std::memset(bitangents, 0, 3 * nv * sizeof(float)); }
    return 0;
}

}  // extern "C"
