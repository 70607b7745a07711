// ORACLE (test infrastructure, never shipped, never on the product path).
//
// CPU restatement of RayTracer/Shapes.h, ThirdParty/AABB_triangle_Moller.h,
// RayTracer/AssetManager.h (data model) and RayTracer/Octtree_Model.h.
// Cited file:line are relative to /root/reference.  PARITY STATUS: pinned against the reference's
// own headers compiled unmodified (oracle/_ref, tests/test_cpu_ref_pin.py): octrees node for node,
// hit ids / t / barycentrics / surface records and all seven analytic shapes bit-identical (sphere
// v within 1 ulp of acos, see tests/ref_pin_cases.py TOLERANT).
#pragma once
#include <chrono>
#include <iostream>
#include <optional>
#include <queue>
#include <string>
#include <unordered_map>
#include <vector>

#include "oracle_pbrt.h"

namespace orc {

// Instrumentation (replaces the racy Hitdata::triangle_intersect_count, Shapes.h:909-911).
struct TraverseCounters {
    uint64_t rays = 0, nodes_visited = 0, tris_tested = 0, leaves_visited = 0, max_queue = 0;
    void add(const TraverseCounters& o) {
        rays += o.rays; nodes_visited += o.nodes_visited; tris_tested += o.tris_tested;
        leaves_visited += o.leaves_visited; max_queue = std::max(max_queue, o.max_queue);
    }
};

// ---- Shapes.h:33-50 ---------------------------------------------------------------------
struct Ray {
    vec3 o, d;
    Ray() = default;
    Ray(vec3 o_, vec3 d_) : o(o_), d(d_) {}
    void Transform(const mat4& M) {
        o = xyz(mul(M, vec4(o.x, o.y, o.z, 1)));
        d = xyz(normalize(mul(M, vec4(d.x, d.y, d.z, 0))));
    }
};

// ---- Shapes.h:53-127 --------------------------------------------------------------------
struct Bounds3 {
    vec3 pmin, pmax;
    Bounds3() = default;
    Bounds3(vec3 a, vec3 b) : pmin(a), pmax(b) {}
    void Transform(const mat4& M) {
        vec3 p[8] = {xyz(mul(M, vec4(pmin.x, pmin.y, pmin.z, 1))), xyz(mul(M, vec4(pmin.x, pmax.y, pmin.z, 1))),
                     xyz(mul(M, vec4(pmin.x, pmax.y, pmax.z, 1))), xyz(mul(M, vec4(pmin.x, pmin.y, pmax.z, 1))),
                     xyz(mul(M, vec4(pmax.x, pmax.y, pmax.z, 1))), xyz(mul(M, vec4(pmax.x, pmin.y, pmax.z, 1))),
                     xyz(mul(M, vec4(pmax.x, pmin.y, pmin.z, 1))), xyz(mul(M, vec4(pmax.x, pmax.y, pmin.z, 1)))};
        // the max side starts at FLT_MIN (smallest positive), not lowest()
        const float big = std::numeric_limits<float>::max(), tiny = std::numeric_limits<float>::min();
        vec2 xm(big, tiny), ym(big, tiny), zm(big, tiny);
        for (int i = 0; i < 8; ++i) {
            xm.x = std::min(xm.x, p[i].x); xm.y = std::max(xm.y, p[i].x);
            ym.x = std::min(ym.x, p[i].y); ym.y = std::max(ym.y, p[i].y);
            zm.x = std::min(zm.x, p[i].z); zm.y = std::max(zm.y, p[i].z);
        }
        pmin = vec3(xm.x, ym.x, zm.x);
        pmax = vec3(xm.y, ym.y, zm.y);
    }
    bool IntersectP(const Ray& ray, float tMax = std::numeric_limits<float>::max(), float* hitt0 = nullptr, float* hitt1 = nullptr) const {
        float min_t = 0, max_t = tMax;
        for (int i = 0; i < 3; ++i) {
            float invRayDir = 1 / ray.d[i];
            float tNear = (pmin[i] - ray.o[i]) * invRayDir;
            float tFar = (pmax[i] - ray.o[i]) * invRayDir;
            if (tNear > tFar) std::swap(tNear, tFar);
            tFar *= 1 + 2 * gamma_n(3);
            min_t = tNear > min_t ? tNear : min_t;
            max_t = tFar < max_t ? tFar : max_t;
            if (min_t > max_t) return false;
        }
        if (hitt0) *hitt0 = min_t;
        if (hitt1) *hitt1 = max_t;
        return true;
    }
};
inline Bounds3 TransformBounds(Bounds3 b, const mat4& M) { b.Transform(M); return b; }

// ---- Shapes.h:144-170 -------------------------------------------------------------------
struct LocalSurfaceInfo {
    float tHit = 0;
    vec3 hitp;
    float u = 0, v = 0;
    vec3 du, dv, n, wo;
    void Transform(const mat4& M) {
        mat3 nt = transpose(upper3(inverse(M)));   // glm::mat3(mat4) takes the upper-left block
        n = normalize(mul(nt, n));
        wo = normalize(mul(nt, wo));
        hitp = xyz(mul(M, vec4(hitp, 1.0f)));
        du = xyz(mul(M, vec4(du, 0)));
        dv = xyz(mul(M, vec4(dv, 0)));
    }
};

// ---- Shapes.h:172-207 -------------------------------------------------------------------
struct Shape {
    std::string name;
    mat4 RenderToObject, ObjectToRender;
    Shape(const std::string& n, const mat4& rigid) : name(n) { SetRigidTransformBase(rigid); }
    virtual ~Shape() = default;
    void SetRigidTransformBase(const mat4& rigid) {
        mat4 perm = mat4::from_cols({1, 0, 0, 0}, {0, 0, 1, 0}, {0, 1, 0, 0}, {0, 0, 0, 1});
        ObjectToRender = mul(rigid, perm);
        RenderToObject = inverse(ObjectToRender);
    }
    virtual void SetRigidTransform(const mat4& rigid) { SetRigidTransformBase(rigid); }
    virtual Bounds3 Bounds() const = 0;
    virtual std::optional<LocalSurfaceInfo> Intersect(const Ray& ray, float tMax = std::numeric_limits<float>::max()) const = 0;
    virtual bool IntersectP(const Ray& ray, float tMax = std::numeric_limits<float>::max()) const = 0;
    virtual float Area() const = 0;
};

constexpr double kPiD = 3.141592653589793238462643383279502884;   // std::numbers::pi

// ---- Shapes.h:209-446 -------------------------------------------------------------------
struct Sphere : Shape {
    struct SphereIntersect { float t; vec3 hitp; vec3 ray_d; float phi; };
    float r, zmin, zmax, thetamin, thetamax, phimax;
    Sphere(const std::string& n, const mat4& rigid, float radius, float zmin_, float zmax_, float phimax_) : Shape(n, rigid) {
        r = radius;
        zmin = clampf(zmin_, -r, r);
        zmax = clampf(zmax_, -r, r);
        thetamin = std::acos(clampf(zmin / r, -1.f, 1.f));
        thetamax = std::acos(clampf(zmax / r, -1.f, 1.f));
        phimax = radians(clampf(phimax_, 0.0f, 360.f));
    }
    float Area() const override { return phimax * r * (zmax - zmin); }
    Bounds3 Bounds() const override { return TransformBounds(Bounds3(vec3(-r, -r, zmin), vec3(r, r, zmax)), ObjectToRender); }
    static float wrap_phi(float y, float x) {
        float phi = std::atan2(y, x);
        if (phi < 0) phi = (float)(phi + 2 * kPiD);   // float += double
        return phi;
    }
    std::optional<SphereIntersect> BasicIntersect(const Ray& ray, float tMax) const {
        vec3 o = xyz(mul(RenderToObject, vec4(ray.o.x, ray.o.y, ray.o.z, 1)));
        vec3 d = xyz(mul(RenderToObject, vec4(ray.d.x, ray.d.y, ray.d.z, 0)));
        float a = d.x * d.x + d.y * d.y + d.z * d.z;
        float b = 2 * (d.x * o.x + d.y * o.y + d.z * o.z);
        float c = o.x * o.x + o.y * o.y + o.z * o.z - r * r;
        vec3 v = o - b / (2 * a) * d;
        float len = length(v);
        float discrim = 4 * a * (r + len) * (r - len);
        if (discrim < 0) return {};
        float rootDiscrim = std::sqrt(discrim);
        float q = (b < 0) ? -.5f * (b - rootDiscrim) : -.5f * (b + rootDiscrim);
        float t0 = q / a, t1 = c / q;
        if (t0 > t1) std::swap(t0, t1);
        if (t0 > tMax || t1 <= 0) return {};
        float tShapeHit = t0;
        if (tShapeHit <= 0) {
            tShapeHit = t1;
            if (tShapeHit > tMax) return {};
        }
        auto refine = [&](float t, vec3& hp, float& phi) {
            hp = o + t * d;
            hp *= r / distance(hp, vec3(0, 0, 0));
            if (hp.x == 0 && hp.y == 0) hp.x = (float)(1e-5 * r);
            phi = wrap_phi(hp.y, hp.x);
        };
        vec3 hitp; float phi;
        refine(tShapeHit, hitp, phi);
        if (hitp.z < zmin || hitp.z > zmax || phi > phimax) {
            if (tShapeHit == t1) return {};
            if (t1 > tMax) return {};
            tShapeHit = t1;
            refine(tShapeHit, hitp, phi);
            if (hitp.z < zmin || hitp.z > zmax || phi > phimax) return {};
        }
        return SphereIntersect{tShapeHit, hitp, normalize(d), phi};
    }
    vec2 PostoAngles(float x, float y, float z) const {
        float theta = std::acos(clampf(z / r, -1.f, 1.f));
        return {theta, wrap_phi(y, x)};
    }
    std::optional<LocalSurfaceInfo> Intersect(const Ray& ray, float tMax = std::numeric_limits<float>::max()) const override {
        auto is = BasicIntersect(ray, tMax);
        if (!is) return {};
        vec3 p = is->hitp;
        LocalSurfaceInfo info;
        info.tHit = is->t;
        info.hitp = p;
        vec2 ang = PostoAngles(p.x, p.y, p.z);
        info.u = clampf(ang.y / phimax, 0.0f, 1.0f);
        info.v = clampf((ang.x - thetamin) / (thetamax - thetamin), 0.0f, 1.0f);
        info.du = normalize(vec3(-phimax * p.y, phimax * p.x, 0));
        {
            float theta = ang.x, phi = ang.y;
            info.dv = normalize((thetamax - thetamin) * vec3(p.z * std::cos(phi), p.z * std::sin(phi), -r * std::sin(theta)));
        }
        info.n = normalize(vec3(2 * p.x, 2 * p.y, 2 * p.z));
        if (dot(info.n, is->ray_d) > 0) info.n = -info.n;
        info.wo = vec3(0, 0, 1);
        info.Transform(ObjectToRender);
        return info;
    }
    bool IntersectP(const Ray& ray, float tMax = std::numeric_limits<float>::max()) const override { return BasicIntersect(ray, tMax).has_value(); }
};

// ---- Shapes.h:448-635 -------------------------------------------------------------------
struct Cylinder : Shape {
    struct CylinderIntersect { float t; vec3 hitp; vec3 ray_d; float phi; };
    float r, min_z, max_z, max_phi;
    Cylinder(const std::string& n, const mat4& rigid, float radius, float zmin, float zmax, float phi_max) : Shape(n, rigid) {
        r = radius; min_z = zmin; max_z = zmax; max_phi = radians(phi_max);
    }
    float Area() const override { return (max_z - min_z) * r * max_phi; }
    Bounds3 Bounds() const override { return TransformBounds(Bounds3(vec3(-r, -r, min_z), vec3(r, r, max_z)), ObjectToRender); }
    std::optional<CylinderIntersect> BasicIntersect(const Ray& ray, float tMax) const {
        vec3 o = xyz(mul(RenderToObject, vec4(ray.o.x, ray.o.y, ray.o.z, 1)));
        vec3 d = xyz(mul(RenderToObject, vec4(ray.d.x, ray.d.y, ray.d.z, 0)));
        float a = d.x * d.x + d.y * d.y;
        float b = 2 * (d.x * o.x + d.y * o.y);
        float c = o.x * o.x + o.y * o.y - r * r;
        float f = b / (2 * a);
        float vx = o.x - f * d.x, vy = o.y - f * d.y;
        float len = std::sqrt(vx * vx + vy * vy);
        float discrim = 4 * a * (r + len) * (r - len);
        if (discrim < 0) return {};
        float rootDiscrim = std::sqrt(discrim);
        float q = (b < 0) ? -.5f * (b - rootDiscrim) : -.5f * (b + rootDiscrim);
        float t0 = q / a, t1 = c / q;
        if (t0 > t1) std::swap(t0, t1);
        if (t0 > tMax || t1 <= 0) return {};
        float tShapeHit = t0;
        if (tShapeHit <= 0) {
            tShapeHit = t1;
            if (tShapeHit > tMax) return {};
        }
        vec3 hit_p = o + tShapeHit * d;
        float phi = Sphere::wrap_phi(hit_p.y, hit_p.x);
        if (hit_p.z < min_z || hit_p.z > max_z || phi > max_phi) {
            if (tShapeHit == t1) return {};
            tShapeHit = t1;
            if (t1 > tMax) return {};
            hit_p = o + tShapeHit * d;
            phi = Sphere::wrap_phi(hit_p.y, hit_p.x);
            if (hit_p.z < min_z || hit_p.z > max_z || phi > max_phi) return {};
        }
        return CylinderIntersect{tShapeHit, hit_p, d, phi};
    }
    std::optional<LocalSurfaceInfo> Intersect(const Ray& ray, float tMax = std::numeric_limits<float>::max()) const override {
        auto is = BasicIntersect(ray, tMax);
        if (!is) return {};
        vec3 p = is->hitp;
        LocalSurfaceInfo info;
        info.tHit = is->t;
        info.hitp = p;
        float phi = Sphere::wrap_phi(p.y, p.x);
        info.u = clampf(phi / max_phi, 0.0f, 1.0f);
        info.v = clampf((p.z - min_z) / (max_z - min_z), 0.0f, 1.0f);
        info.du = normalize(vec3(-max_phi * p.y, max_phi * p.x, 0));
        info.dv = normalize(vec3(0, 0, max_z - min_z));
        info.n = normalize(vec3(2 * p.x, 2 * p.y, 0));
        if (dot(info.n, is->ray_d) > 0) info.n = -info.n;
        info.wo = vec3(0, 0, 1);
        info.Transform(ObjectToRender);
        return info;
    }
    bool IntersectP(const Ray& ray, float tMax = std::numeric_limits<float>::max()) const override { return BasicIntersect(ray, tMax).has_value(); }
};

// ---- Shapes.h:637-779 -------------------------------------------------------------------
struct Disk : Shape {
    struct DiskIntersect { float t; vec3 hitp; vec3 ray_d; float phi; };
    float inner_r, outer_r, h, phimax;
    Disk(const std::string& n, const mat4& rigid, float height, float in_r, float out_r, float phimax_) : Shape(n, rigid) {
        inner_r = in_r; outer_r = out_r; h = height; phimax = radians(phimax_);
    }
    float Area() const override { return phimax * .5f * (outer_r * outer_r - inner_r * inner_r); }
    Bounds3 Bounds() const override { return TransformBounds(Bounds3(vec3(-outer_r, -outer_r, h), vec3(outer_r, outer_r, h)), ObjectToRender); }
    std::optional<DiskIntersect> BasicIntersect(const Ray& ray, float tMax) const {
        vec3 o = xyz(mul(RenderToObject, vec4(ray.o.x, ray.o.y, ray.o.z, 1)));
        vec3 d = xyz(mul(RenderToObject, vec4(ray.d.x, ray.d.y, ray.d.z, 0)));
        float t0 = (h - o.z) / d.z;                 // computed before the d.z == 0 check (Shapes.h:691-696)
        if (t0 <= 0 || t0 >= tMax) return {};
        if (d.z == 0) return {};
        vec3 phit = o + t0 * d;
        float dist2 = phit.x * phit.x + phit.y * phit.y;
        if (dist2 > outer_r * outer_r || dist2 < inner_r * inner_r) return {};
        float phi = Sphere::wrap_phi(phit.y, phit.x);
        if (phi > phimax) return {};
        return DiskIntersect{t0, phit, normalize(d), phi};
    }
    std::optional<LocalSurfaceInfo> Intersect(const Ray& ray, float tMax = std::numeric_limits<float>::max()) const override {
        auto is = BasicIntersect(ray, tMax);
        if (!is) return {};
        vec3 p = is->hitp;
        LocalSurfaceInfo info;
        info.tHit = is->t;
        info.hitp = p;
        float phi = Sphere::wrap_phi(p.y, p.x);
        info.u = clampf(phi / phimax, 0.0f, 1.0f);
        info.v = clampf((outer_r - std::sqrt(p.x * p.x + p.y * p.y)) / (outer_r - inner_r), 0.0f, 1.0f);
        info.du = normalize(vec3(-phimax * p.y, phimax * p.x, 0));
        info.dv = normalize(vec3(p.x, p.y, 0) * (inner_r - outer_r) / std::sqrt(p.x * p.x + p.y * p.y));
        info.n = vec3(0, 0, 1);
        if (dot(info.n, is->ray_d) > 0) info.n = -info.n;
        info.wo = vec3(0, 0, 1);
        info.Transform(ObjectToRender);
        return info;
    }
    bool IntersectP(const Ray& ray, float tMax = std::numeric_limits<float>::max()) const override { return BasicIntersect(ray, tMax).has_value(); }
};

// ---- Shapes.h:781-905 -------------------------------------------------------------------
struct TriangleSimple : Shape {
    struct TriangleIntersect { float t; vec3 hitp; vec3 ray_d; float B, Y; };
    vec3 p1, p2, p3;
    TriangleSimple(const std::string& n, const mat4& rigid, vec3 a, vec3 b, vec3 c) : Shape(n, rigid), p1(a), p2(b), p3(c) {}
    float Area() const override { return 0.5f * length(cross(p2 - p1, p3 - p1)); }
    Bounds3 Bounds() const override {
        return TransformBounds(Bounds3(vec3(gmin(gmin(p1.x, p2.x), p3.x), gmin(gmin(p1.y, p2.y), p3.y), gmin(gmin(p1.z, p2.z), p3.z)),
                                       vec3(gmax(gmax(p1.x, p2.x), p3.x), gmax(gmax(p1.y, p2.y), p3.y), gmax(gmax(p1.z, p2.z), p3.z))),
                               ObjectToRender);
    }
    std::optional<TriangleIntersect> BasicIntersect(const Ray& ray, float tMax) const {
        vec3 orig = xyz(mul(RenderToObject, vec4(ray.o.x, ray.o.y, ray.o.z, 1)));
        vec3 dir = xyz(mul(RenderToObject, vec4(ray.d.x, ray.d.y, ray.d.z, 0)));
        float a = p1.x - p2.x, b = p1.y - p2.y, c = p1.z - p2.z;
        float d = p1.x - p3.x, e = p1.y - p3.y, f = p1.z - p3.z;
        float g = dir.x, h = dir.y, i = dir.z;
        float j = p1.x - orig.x, k = p1.y - orig.y, l = p1.z - orig.z;
        float M = a * (e * i - h * f) + b * (g * f - d * i) + c * (d * h - e * g);
        float t = -(f * (a * k - j * b) + e * (j * c - a * l) + d * (b * l - k * c)) / M;
        if (t < 0 || t >= tMax) return {};
        float Y = (i * (a * k - j * b) + h * (j * c - a * l) + g * (b * l - k * c)) / M;
        if (Y < 0 || Y > 1) return {};
        float B = (j * (e * i - h * f) + k * (g * f - d * i) + l * (d * h - e * g)) / M;
        if (B < 0 || B > 1 - Y) return {};
        vec3 hitp = orig + t * dir;
        return TriangleIntersect{t, hitp, normalize(dir), B, Y};
    }
    std::optional<LocalSurfaceInfo> Intersect(const Ray& ray, float tMax = std::numeric_limits<float>::max()) const override {
        auto is = BasicIntersect(ray, tMax);
        if (!is) return {};
        LocalSurfaceInfo info;
        info.tHit = is->t;
        info.hitp = is->hitp;
        info.u = clampf(is->B, 0.0f, 1.0f);
        info.v = clampf(is->Y, 0.0f, 1.0f);
        info.du = normalize(p2 - p1);
        info.dv = normalize(p3 - p1);
        info.n = normalize(cross(p3 - p1, p2 - p1));
        if (dot(info.n, is->ray_d) > 0) info.n = -info.n;
        info.wo = vec3(0, 0, 1);
        info.Transform(ObjectToRender);
        return info;
    }
    bool IntersectP(const Ray& ray, float tMax = std::numeric_limits<float>::max()) const override { return BasicIntersect(ray, tMax).has_value(); }
};

// ---- AssetManager.h:20-47,65 ------------------------------------------------------------
struct MeshCache {
    struct Mesh {
        std::vector<vec3> positions, normals;
        std::vector<vec2> texcoords;
        std::vector<vec3> tangents, bitangents;
        std::vector<unsigned int> indices;
    };
    struct Model { std::vector<Mesh> meshes; std::string mesh_name; };
    static std::unordered_map<std::string, Model>& modelCache() {
        static std::unordered_map<std::string, Model> c;
        return c;
    }
};

// When true, Triangle::BasicIntersect also performs the reference's per-call overheads (two
// unordered_map<string> lookups, a chrono::now(), Shapes.h:1103-1111) so the "cpu_faithful"
// baseline of BASELINE.md can be timed.  Results are identical either way.
extern bool g_faithful_overheads;
extern thread_local TraverseCounters* tl_counters;

// ---- Shapes.h:913-1270 ------------------------------------------------------------------
struct TriangleVertexAvailable {      // Triangle::vertex_available, Shapes.h:917-924
    bool texcoords = true, normals = true, tangents = true, bitangents = true, precomputed_worldtransform = false;
};
struct Triangle : Shape {
    using vertex_available = TriangleVertexAvailable;
    struct TriangleIntersect { float b0, b1, b2, t; vec3 rayd; };
    std::string model_name;
    int mesh_id, tri_id;
    vertex_available available_info;
    const MeshCache::Mesh* mesh_fast = nullptr;   // resolved once; the faithful path re-resolves per call

    Triangle(const std::string& n, const mat4& rigid, const std::string& model, int mesh, int tri, vertex_available avail = vertex_available())
        : Shape(n, rigid), model_name(model), mesh_id(mesh), tri_id(tri), available_info(avail) {
        auto it = MeshCache::modelCache().find(model_name);
        if (it != MeshCache::modelCache().end()) mesh_fast = &it->second.meshes[mesh_id];
    }
    const MeshCache::Mesh* mesh_lookup() const {
        if (!g_faithful_overheads) return mesh_fast;
        auto& cache = MeshCache::modelCache();
        if (cache.find(model_name) == cache.end()) return nullptr;       // Shapes.h:1103
        volatile auto t0 = std::chrono::high_resolution_clock::now().time_since_epoch().count();   // Timer, :1109-1110
        (void)t0;
        return &cache[model_name].meshes[mesh_id];                        // :1111
    }
    float Area() const override {
        const auto* mesh = mesh_fast;
        if (!mesh) return 0.0f;
        vec3 p0 = mesh->positions[mesh->indices[3 * tri_id]], p1 = mesh->positions[mesh->indices[3 * tri_id + 1]], p2 = mesh->positions[mesh->indices[3 * tri_id + 2]];
        return 0.5f * length(cross(p1 - p0, p2 - p0));
    }
    Bounds3 Bounds() const override {
        const auto* mesh = mesh_fast;
        if (!mesh) return Bounds3(vec3(0, 0, 0), vec3(0, 0, 0));
        vec3 p0 = mesh->positions[mesh->indices[3 * tri_id]], p1 = mesh->positions[mesh->indices[3 * tri_id + 1]], p2 = mesh->positions[mesh->indices[3 * tri_id + 2]];
        return TransformBounds(Bounds3(vec3(gmin(gmin(p0.x, p1.x), p2.x), gmin(gmin(p0.y, p1.y), p2.y), gmin(gmin(p0.z, p1.z), p2.z)),
                                       vec3(gmax(gmax(p0.x, p1.x), p2.x), gmax(gmax(p0.y, p1.y), p2.y), gmax(gmax(p0.z, p1.z), p2.z))),
                               ObjectToRender);
    }

    // Shapes.h:1101-1260 -- pbrt-v4 watertight ray/triangle test
    std::optional<TriangleIntersect> BasicIntersect(const Ray& ray, float tMax = std::numeric_limits<float>::max()) const {
        const MeshCache::Mesh* mesh = mesh_lookup();
        if (!mesh) return {};
        if (tl_counters) tl_counters->tris_tested++;
        vec3 p0 = mesh->positions[mesh->indices[3 * tri_id]];
        vec3 p1 = mesh->positions[mesh->indices[3 * tri_id + 1]];
        vec3 p2 = mesh->positions[mesh->indices[3 * tri_id + 2]];
        vec3 p0_w, p1_w, p2_w;
        if (!available_info.precomputed_worldtransform) {
            p0_w = xyz(mul(ObjectToRender, vec4(p0.x, p0.y, p0.z, 1)));
            p1_w = xyz(mul(ObjectToRender, vec4(p1.x, p1.y, p1.z, 1)));
            p2_w = xyz(mul(ObjectToRender, vec4(p2.x, p2.y, p2.z, 1)));
        } else { p0_w = p0; p1_w = p1; p2_w = p2; }

        // degenerate: std::pow(length(cross), 2) == 0  <=>  length == 0
        if (std::pow(length(cross(p2_w - p0_w, p1_w - p0_w)), 2) == 0) return {};

        vec3 p0t = p0_w - ray.o, p1t = p1_w - ray.o, p2t = p2_w - ray.o;
        vec3 ad(std::abs(ray.d.x), std::abs(ray.d.y), std::abs(ray.d.z));
        int kz = MaxComponentIndex(ad);
        int kx = kz + 1; if (kx == 3) kx = 0;
        int ky = kx + 1; if (ky == 3) ky = 0;
        vec3 d(ray.d[kx], ray.d[ky], ray.d[kz]);
        p0t = vec3(p0t[kx], p0t[ky], p0t[kz]);
        p1t = vec3(p1t[kx], p1t[ky], p1t[kz]);
        p2t = vec3(p2t[kx], p2t[ky], p2t[kz]);

        float Sx = -d.x / d.z, Sy = -d.y / d.z, Sz = 1 / d.z;
        p0t.x += Sx * p0t.z; p0t.y += Sy * p0t.z;
        p1t.x += Sx * p1t.z; p1t.y += Sy * p1t.z;
        p2t.x += Sx * p2t.z; p2t.y += Sy * p2t.z;

        float e0 = DifferenceOfProducts(p1t.x, p2t.y, p1t.y, p2t.x);
        float e1 = DifferenceOfProducts(p2t.x, p0t.y, p2t.y, p0t.x);
        float e2 = DifferenceOfProducts(p0t.x, p1t.y, p0t.y, p1t.x);
        if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {      // double fallback, Shapes.h:1174-1184
            double p2txp1ty = (double)p2t.x * (double)p1t.y, p2typ1tx = (double)p2t.y * (double)p1t.x;
            e0 = (float)(p2typ1tx - p2txp1ty);
            double p0txp2ty = (double)p0t.x * (double)p2t.y, p0typ2tx = (double)p0t.y * (double)p2t.x;
            e1 = (float)(p0typ2tx - p0txp2ty);
            double p1txp0ty = (double)p1t.x * (double)p0t.y, p1typ0tx = (double)p1t.y * (double)p0t.x;
            e2 = (float)(p1typ0tx - p1txp0ty);
        }
        if ((e0 < 0 || e1 < 0 || e2 < 0) && (e0 > 0 || e1 > 0 || e2 > 0)) return {};
        float det = e0 + e1 + e2;
        if (det == 0) return {};

        p0t.z *= Sz; p1t.z *= Sz; p2t.z *= Sz;
        float tScaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
        if (det < 0 && (tScaled >= 0 || tScaled < tMax * det)) return {};
        else if (det > 0 && (tScaled <= 0 || tScaled > tMax * det)) return {};

        float invDet = 1 / det;
        float b0 = e0 * invDet, b1 = e1 * invDet, b2 = e2 * invDet;
        float t = tScaled * invDet;
        if (std::isnan(t)) return {};

        float maxZt = MaxComponentValue(vec3(std::abs(p0t.z), std::abs(p1t.z), std::abs(p2t.z)));
        float deltaZ = gamma_n(3) * maxZt;
        float maxXt = MaxComponentValue(vec3(std::abs(p0t.x), std::abs(p1t.x), std::abs(p2t.x)));
        float maxYt = MaxComponentValue(vec3(std::abs(p0t.y), std::abs(p1t.y), std::abs(p2t.y)));
        float deltaX = gamma_n(5) * (maxXt + maxZt);
        float deltaY = gamma_n(5) * (maxYt + maxZt);
        float deltaE = 2 * (gamma_n(2) * maxXt * maxYt + deltaY * maxXt + deltaX * maxYt);
        float maxE = MaxComponentValue(vec3(std::abs(e0), std::abs(e1), std::abs(e2)));
        float deltaT = 3 * (gamma_n(3) * maxE * maxZt + deltaE * maxZt + deltaZ * maxE) * std::abs(invDet);
        if (t <= deltaT) return {};
        return TriangleIntersect{b0, b1, b2, t, normalize(ray.d)};
    }

    // Shapes.h:982-1083
    std::optional<LocalSurfaceInfo> CalculateLocalSurface(const TriangleIntersect& isect) const {
        const MeshCache::Mesh* meshp = mesh_fast;
        if (!meshp) return {};
        const MeshCache::Mesh& mesh = *meshp;
        unsigned i0 = mesh.indices[3 * tri_id], i1 = mesh.indices[3 * tri_id + 1], i2 = mesh.indices[3 * tri_id + 2];
        vec3 p0 = mesh.positions[i0], p1 = mesh.positions[i1], p2 = mesh.positions[i2];
        // NB: the reference transforms unconditionally here, even for precomputed world positions (Shapes.h:995-997)
        vec3 p0_w = xyz(mul(ObjectToRender, vec4(p0.x, p0.y, p0.z, 1)));
        vec3 p1_w = xyz(mul(ObjectToRender, vec4(p1.x, p1.y, p1.z, 1)));
        vec3 p2_w = xyz(mul(ObjectToRender, vec4(p2.x, p2.y, p2.z, 1)));
        vec2 uv[3] = {vec2(0, 0), vec2(1, 0), vec2(0, 1)};
        vec2 duv02 = uv[0] - uv[2], duv12 = uv[1] - uv[2];
        vec3 dp02 = p0_w - p2_w, dp12 = p1_w - p2_w;
        float determinant = duv02.x * duv12.y - duv02.y * duv12.x;
        vec3 dpdu, dpdv;
        bool degenerateUV = std::abs(determinant) < 1e-9f;
        if (!degenerateUV) {
            float invdet = 1 / determinant;
            dpdu = (duv12.y * dp02 - duv02.y * dp12) * invdet;
            dpdv = (duv02.x * dp12 - duv12.x * dp02) * invdet;
        }
        if (degenerateUV || std::pow(length(cross(dpdu, dpdv)), 2) == 0) {
            vec3 ng = cross(p2_w - p0_w, p1_w - p0_w);
            if (std::pow(length(ng), 2) == 0) {
                vec3 a = p2_w - p0_w, b = p1_w - p0_w;
                dvec3 c = cross(dvec3{a.x, a.y, a.z}, dvec3{b.x, b.y, b.z});
                ng = vec3((float)c.x, (float)c.y, (float)c.z);
            }
            ng = normalize(ng);
            float sign = std::copysign(float(1), ng.z);
            float a = -1 / (sign + ng.z);
            float b = ng.x * ng.y * a;
            dpdu = vec3((float)(1 + sign * std::pow(ng.x, 2) * a), sign * b, -sign * ng.x);
            dpdv = vec3(b, (float)(sign + std::pow(ng.y, 2) * a), -ng.y);
        }
        LocalSurfaceInfo info;
        info.hitp = p0_w * isect.b0 + p1_w * isect.b1 + p2_w * isect.b2;
        vec2 hit_uv;
        if (available_info.texcoords) hit_uv = mesh.texcoords[i0] * isect.b0 + mesh.texcoords[i1] * isect.b1 + mesh.texcoords[i2] * isect.b2;
        else hit_uv = uv[0] * isect.b0 + uv[1] * isect.b1 + uv[2] * isect.b2;
        info.u = clampf(hit_uv.x, 0.0f, 1.0f);
        info.v = clampf(hit_uv.y, 0.0f, 1.0f);
        if (available_info.tangents) info.du = normalize(mesh.tangents[i0] * isect.b0 + mesh.tangents[i1] * isect.b1 + mesh.tangents[i2] * isect.b2);
        else info.du = dpdu;
        if (available_info.bitangents) info.dv = normalize(mesh.bitangents[i0] * isect.b0 + mesh.bitangents[i1] * isect.b1 + mesh.bitangents[i2] * isect.b2);
        else info.dv = dpdv;
        if (available_info.normals) info.n = normalize(mesh.normals[i0] * isect.b0 + mesh.normals[i1] * isect.b1 + mesh.normals[i2] * isect.b2);
        else info.n = normalize(cross(dp02, dp12));
        if (dot(info.n, isect.rayd) > 0) info.n = -info.n;
        info.wo = vec3(0, 0, 1);
        return info;   // tHit is never written by the reference
    }
    std::optional<LocalSurfaceInfo> Intersect(const Ray& ray, float tMax = std::numeric_limits<float>::max()) const override {
        auto is = BasicIntersect(ray, tMax);
        if (!is) return {};
        return CalculateLocalSurface(*is);
    }
    bool IntersectP(const Ray& ray, float tMax = std::numeric_limits<float>::max()) const override { return BasicIntersect(ray, tMax).has_value(); }
};

// ---- Shapes.h:1273-1491 -----------------------------------------------------------------
struct TriModel : Shape {
    struct TriModelIntersect { Triangle::TriangleIntersect tri_isect; int mesh_id, tri_id; };
    std::string model_name;
    Triangle::vertex_available avail_info;
    std::vector<std::vector<Triangle>> triangles;
    Bounds3 precomputed_bounds;
    bool enable_cull_back_face;
    std::vector<std::vector<bool>> back_facing;
    bool mesh_has_precomputed_worldposition;

    TriModel(const std::string& n, const mat4& rigid, const std::string& model, bool cull, bool precomp_world, Triangle::vertex_available avail)
        : Shape(n, rigid), model_name(model), avail_info(avail), enable_cull_back_face(cull), mesh_has_precomputed_worldposition(precomp_world) {
        MeshCache::Model& m = MeshCache::modelCache()[model_name];
        const float big = std::numeric_limits<float>::max(), tiny = std::numeric_limits<float>::min();
        vec3 mn(big, big, big), mx(tiny, tiny, tiny);
        for (auto& mesh : m.meshes)
            for (const vec3& p : mesh.positions) {
                mn.x = std::min(mn.x, p.x); mn.y = std::min(mn.y, p.y); mn.z = std::min(mn.z, p.z);
                mx.x = std::max(mx.x, p.x); mx.y = std::max(mx.y, p.y); mx.z = std::max(mx.z, p.z);
            }
        precomputed_bounds = Bounds3(mn, mx);
        triangles.resize(m.meshes.size());
        for (size_t mi = 0; mi < m.meshes.size(); ++mi) {
            size_t nt = m.meshes[mi].indices.size() / 3;
            triangles[mi].reserve(nt);
            for (size_t ti = 0; ti < nt; ++ti) triangles[mi].emplace_back("tri", rigid, model_name, (int)mi, (int)ti, avail_info);
        }
    }
    void SetRigidTransform(const mat4& rigid) override {
        for (auto& v : triangles) for (auto& t : v) t.SetRigidTransform(rigid);
        SetRigidTransformBase(rigid);
    }
    void EnableBackface(bool v) { enable_cull_back_face = v; }
    void ComputeBackFace(vec3 world_pos_look, bool enable) {     // Shapes.h:1339-1380
        enable_cull_back_face = enable;
        if (!enable_cull_back_face) return;
        MeshCache::Model& m = MeshCache::modelCache()[model_name];
        vec3 look = normalize(world_pos_look);
        back_facing.clear();
        back_facing.resize(m.meshes.size());
        for (size_t mi = 0; mi < m.meshes.size(); ++mi) {
            auto& mesh = m.meshes[mi];
            size_t nt = mesh.indices.size() / 3;
            back_facing[mi].reserve(nt);
            for (size_t ti = 0; ti < nt; ++ti) {
                vec3 n1 = mesh.normals[mesh.indices[3 * ti]], n2 = mesh.normals[mesh.indices[3 * ti + 1]], n3 = mesh.normals[mesh.indices[3 * ti + 2]];
                vec3 N = normalize((n1 + n2 + n3) / 3.0f);
                if (!mesh_has_precomputed_worldposition) {
                    mat3 nt3 = transpose(upper3(inverse(ObjectToRender)));
                    N = normalize(mul(nt3, N));
                }
                back_facing[mi].push_back(dot(look, N) > 0);
            }
        }
    }
    float Area() const override {
        return (precomputed_bounds.pmax.x - precomputed_bounds.pmin.x) * (precomputed_bounds.pmax.y - precomputed_bounds.pmin.y) *
               (precomputed_bounds.pmax.z - precomputed_bounds.pmin.z);
    }
    Bounds3 Bounds() const override {
        if (!mesh_has_precomputed_worldposition) return TransformBounds(precomputed_bounds, ObjectToRender);
        return precomputed_bounds;
    }
    // brute force closest hit, Shapes.h:1414-1471 (calls the triangle test twice per hit)
    std::optional<TriModelIntersect> BasicIntersect(const Ray& ray, float tMax = std::numeric_limits<float>::max()) const {
        Bounds3 bb = precomputed_bounds;
        if (!mesh_has_precomputed_worldposition) bb = TransformBounds(precomputed_bounds, ObjectToRender);
        if (!bb.IntersectP(ray, tMax)) return {};
        float t_min = std::numeric_limits<float>::max();
        int id_mesh = 0, id_tri = 0;
        std::optional<Triangle::TriangleIntersect> cur;
        for (size_t mi = 0; mi < triangles.size(); ++mi)
            for (size_t ti = 0; ti < triangles[mi].size(); ++ti) {
                if (enable_cull_back_face && !back_facing.empty() && back_facing[mi][ti]) continue;
                if (triangles[mi][ti].BasicIntersect(ray, tMax).has_value()) {
                    auto is = triangles[mi][ti].BasicIntersect(ray, tMax);
                    if (is.has_value() && is->t < t_min) { id_mesh = (int)mi; id_tri = (int)ti; t_min = is->t; cur = is; }
                }
            }
        if (t_min == std::numeric_limits<float>::max()) return {};
        return TriModelIntersect{*cur, id_mesh, id_tri};
    }
    std::optional<LocalSurfaceInfo> Intersect(const Ray& ray, float tMax = std::numeric_limits<float>::max()) const override {
        auto is = BasicIntersect(ray, tMax);
        if (!is) return {};
        return triangles[is->mesh_id][is->tri_id].CalculateLocalSurface(is->tri_isect);
    }
    bool IntersectP(const Ray& ray, float tMax = std::numeric_limits<float>::max()) const override { return BasicIntersect(ray, tMax).has_value(); }
};

// ---- ThirdParty/AABB_triangle_Moller.h:185-476 ---------------------------------------------
namespace moller {
inline void FindMinMax(float x0, float x1, float x2, float& mn, float& mx) {
    mn = mx = x0;
    if (x1 < mn) mn = x1;
    if (x1 > mx) mx = x1;
    if (x2 < mn) mn = x2;
    if (x2 > mx) mx = x2;
}
inline int planeBoxOverlap(vec3 normal, vec3 vert, vec3 maxbox) {
    vec3 vmin, vmax;
    for (int q = 0; q <= 2; q++) {
        float v = vert[q];
        if (normal[q] > 0.0f) { vmin[q] = -maxbox[q] - v; vmax[q] = maxbox[q] - v; }
        else { vmin[q] = maxbox[q] - v; vmax[q] = -maxbox[q] - v; }
    }
    if (dot(normal, vmin) > 0.0f) return 0;
    if (dot(normal, vmax) >= 0.0f) return 1;
    return 0;
}
// One edge-axis SAT test: p = a*u - b*v at two vertices against rad.  `reject` false reproduces the
// reference's AxisTest_Z0, whose both branches return true (AABB_triangle_Moller.h:334-345).
inline bool axis(float pa, float pb, float rad, bool reject = true) {
    float mn, mx;
    if (pa < pb) { mn = pa; mx = pb; } else { mn = pb; mx = pa; }
    if (mn > rad || mx < -rad) return !reject;
    return true;
}
inline int triBoxOverlap(vec3 boxcenter, vec3 h, const vec3 tv[3]) {
    vec3 v0 = tv[0] - boxcenter, v1 = tv[1] - boxcenter, v2 = tv[2] - boxcenter;
    vec3 e0 = v1 - v0, e1 = v2 - v1, e2 = v0 - v2;
    float fex, fey, fez;
    // X01(a,b,fa,fb): p0 = a*v0.y - b*v0.z, p2 = a*v2.y - b*v2.z, rad = fa*h.y + fb*h.z
    auto X01 = [&](float a, float b, float fa, float fb) { return axis(a * v0.y - b * v0.z, a * v2.y - b * v2.z, fa * h.y + fb * h.z); };
    auto X2 = [&](float a, float b, float fa, float fb) { return axis(a * v0.y - b * v0.z, a * v1.y - b * v1.z, fa * h.y + fb * h.z); };
    auto Y02 = [&](float a, float b, float fa, float fb) { return axis(-a * v0.x + b * v0.z, -a * v2.x + b * v2.z, fa * h.x + fb * h.z); };
    auto Y1 = [&](float a, float b, float fa, float fb) { return axis(-a * v0.x + b * v0.z, -a * v1.x + b * v1.z, fa * h.x + fb * h.z); };
    auto Z0 = [&](float a, float b, float fa, float fb) { return axis(a * v0.x - b * v0.y, a * v1.x - b * v1.y, fa * h.x + fb * h.y, false); };
    // Z12 orders (p2 < p1): min/max are symmetric so axis() is equivalent
    auto Z12 = [&](float a, float b, float fa, float fb) { return axis(a * v2.x - b * v2.y, a * v1.x - b * v1.y, fa * h.x + fb * h.y); };

    fex = std::fabs(e0.x); fey = std::fabs(e0.y); fez = std::fabs(e0.z);
    if (!X01(e0.z, e0.y, fez, fey)) return 0;
    if (!Y02(e0.z, e0.x, fez, fex)) return 0;
    if (!Z12(e0.y, e0.x, fey, fex)) return 0;
    fex = std::fabs(e1.x); fey = std::fabs(e1.y); fez = std::fabs(e1.z);
    if (!X01(e1.z, e1.y, fez, fey)) return 0;
    if (!Y02(e1.z, e1.x, fez, fex)) return 0;
    if (!Z0(e1.y, e1.x, fey, fex)) return 0;
    fex = std::fabs(e2.x); fey = std::fabs(e2.y); fez = std::fabs(e2.z);
    if (!X2(e2.z, e2.y, fez, fey)) return 0;
    if (!Y1(e2.z, e2.x, fez, fex)) return 0;
    if (!Z12(e2.y, e2.x, fey, fex)) return 0;

    float mn, mx;
    FindMinMax(v0.x, v1.x, v2.x, mn, mx);
    if (mn > h.x || mx < -h.x) return 0;
    FindMinMax(v0.y, v1.y, v2.y, mn, mx);
    if (mn > h.y || mx < -h.y) return 0;
    FindMinMax(v0.z, v1.z, v2.z, mn, mx);
    if (mn > h.z || mx < -h.z) return 0;
    vec3 normal = cross(e0, e1);
    if (!planeBoxOverlap(normal, v0, h)) return 0;
    return 1;
}
}  // namespace moller

// ---- Octtree_Model.h ------------------------------------------------------------------------
struct Octtree_Model {
    struct index_info { int mesh_id, tri_id; };
    struct node {
        Bounds3 bounds;
        std::vector<index_info> triangle_info;
        bool leaf = true;
        int parent_id = -1;
        std::vector<int> child_id;
    };
    static const int TRIANGLE_CAPACITY = 40;      // Octtree_Model.h:388
    std::vector<node> octtree;
    TriModel& model;
    explicit Octtree_Model(TriModel& m) : model(m) {}

    void CreateOcttree() {                        // :33-63
        octtree.clear();
        octtree.reserve(10000);
        node root;
        root.bounds = model.Bounds();
        root.leaf = true;
        root.parent_id = 0;
        octtree.push_back(root);
        for (int mi = 0; mi < (int)model.triangles.size(); ++mi)
            for (int ti = 0; ti < (int)model.triangles[mi].size(); ++ti) AddTriangle({mi, ti});
    }
    void tri_world(index_info info, vec3 p[3]) const {
        const MeshCache::Mesh& mesh = MeshCache::modelCache()[model.model_name].meshes[info.mesh_id];
        for (int k = 0; k < 3; ++k) {
            p[k] = mesh.positions[mesh.indices[3 * info.tri_id + k]];
            if (!model.mesh_has_precomputed_worldposition) p[k] = xyz(mul(model.ObjectToRender, vec4(p[k].x, p[k].y, p[k].z, 1)));
        }
    }
    static bool tri_boundsIntersection(const vec3 p[3], const Bounds3& b) {   // :361-366
        vec3 half_d = vec3(b.pmax.x - b.pmin.x, b.pmax.y - b.pmin.y, b.pmax.z - b.pmin.z) / 2.0f;
        vec3 C = b.pmin + half_d;
        return moller::triBoxOverlap(C, half_d, p) != 0;
    }
    void AddTriangle(index_info info) {           // :180-277
        vec3 p[3];
        tri_world(info, p);
        std::queue<int> queue;
        queue.push(0);
        while (!queue.empty()) {
            int cur = queue.front();
            queue.pop();
            if (tri_boundsIntersection(p, octtree[cur].bounds)) {
                if (octtree[cur].leaf) {
                    octtree[cur].triangle_info.push_back(info);
                    if ((int)octtree[cur].triangle_info.size() >= TRIANGLE_CAPACITY) Split(cur);
                } else {
                    for (int i = 0; i < 8; ++i) queue.push(octtree[cur].child_id[i]);
                }
            }
        }
    }
    void Split(int split_id) {                    // :279-358
        Bounds3 B = octtree[split_id].bounds;
        float padding = 0.01f;
        vec3 hd = vec3(B.pmax.x - B.pmin.x, B.pmax.y - B.pmin.y, B.pmax.z - B.pmin.z) / 2.0f;
        vec3 C = B.pmin + hd;
        hd += vec3(padding, padding, padding);
        Bounds3 kids[8] = {
            Bounds3(C + vec3(-hd.x, 0, -hd.z), C + vec3(0, hd.y, 0)),        // top front left
            Bounds3(C + vec3(0, 0, -hd.z), C + vec3(hd.x, hd.y, 0)),         // top front right
            Bounds3(C + vec3(-hd.x, 0, 0), C + vec3(0, hd.y, hd.z)),         // top back left
            Bounds3(C + vec3(0, 0, 0), C + vec3(hd.x, hd.y, hd.z)),          // top back right
            Bounds3(C + vec3(-hd.x, -hd.y, -hd.z), C + vec3(0, 0, 0)),       // bottom front left
            Bounds3(C + vec3(0, -hd.y, -hd.z), C + vec3(hd.x, 0, 0)),        // bottom front right
            Bounds3(C + vec3(-hd.x, -hd.y, 0), C + vec3(0, 0, hd.z)),        // bottom back left
            Bounds3(C + vec3(0, -hd.y, 0), C + vec3(hd.x, 0, hd.z))};        // bottom back right
        std::vector<node> nodes(8);
        for (int n = 0; n < 8; ++n) nodes[n].bounds = kids[n];
        for (size_t i = 0; i < octtree[split_id].triangle_info.size(); ++i) {
            index_info ci = octtree[split_id].triangle_info[i];
            vec3 p[3];
            tri_world(ci, p);
            for (int n = 0; n < 8; ++n)
                if (tri_boundsIntersection(p, nodes[n].bounds)) nodes[n].triangle_info.push_back(ci);
        }
        size_t count = octtree[split_id].triangle_info.size();
        for (int n = 0; n < 8; ++n)
            if (nodes[n].triangle_info.size() == count) return;      // abort: a child got everything
        octtree[split_id].child_id.clear();
        octtree[split_id].child_id.reserve(8);
        for (int n = 0; n < 8; ++n) {
            nodes[n].parent_id = split_id;
            octtree.push_back(nodes[n]);
            octtree[split_id].child_id.push_back((int)octtree.size() - 1);
        }
        octtree[split_id].triangle_info.clear();
        octtree[split_id].leaf = false;
    }

    struct HitRecord { bool found = false; index_info info{-1, -1}; Triangle::TriangleIntersect isect{}; };
    // :66-127 up to (not including) CalculateLocalSurface; BFS with shrinking tMax, strict '<'
    HitRecord TraverseClosest(const Ray& ray, float tMax0 = std::numeric_limits<float>::max()) const {
        HitRecord rec;
        float tMax = tMax0;
        if (tl_counters) tl_counters->rays++;
        std::queue<int> queue;
        queue.push(0);
        while (!queue.empty()) {
            if (tl_counters) tl_counters->max_queue = std::max<uint64_t>(tl_counters->max_queue, queue.size());
            int cur = queue.front();
            queue.pop();
            if (tl_counters) tl_counters->nodes_visited++;
            const node& nd = octtree[cur];
            if (nd.bounds.IntersectP(ray, tMax)) {
                if (nd.leaf) {
                    if (tl_counters) tl_counters->leaves_visited++;
                    for (size_t t = 0; t < nd.triangle_info.size(); ++t) {
                        index_info li = nd.triangle_info[t];
                        if (model.enable_cull_back_face && !model.back_facing.empty() && model.back_facing[li.mesh_id][li.tri_id]) continue;
                        auto is = model.triangles[li.mesh_id][li.tri_id].BasicIntersect(ray, tMax);
                        if (is.has_value() && is->t < tMax) {
                            rec.found = true;
                            tMax = is->t;
                            rec.isect = *is;
                            rec.info = li;
                        }
                    }
                } else {
                    for (int i = 0; i < 8; ++i) queue.push(nd.child_id[i]);
                }
            }
        }
        return rec;
    }
    std::optional<LocalSurfaceInfo> Traverse(Ray& ray) const {
        HitRecord rec = TraverseClosest(ray);
        if (!rec.found) return {};
        return model.triangles[rec.info.mesh_id][rec.info.tri_id].CalculateLocalSurface(rec.isect);
    }
    // Tier B (no reference counterpart): occlusion query with a FIXED tMax -- every triangle test is
    // independent of visit order, so any traversal order gives the same boolean.
    bool TraverseAny(const Ray& ray, float tMax) const {
        if (tl_counters) tl_counters->rays++;
        std::queue<int> queue;
        queue.push(0);
        while (!queue.empty()) {
            int cur = queue.front();
            queue.pop();
            if (tl_counters) tl_counters->nodes_visited++;
            const node& nd = octtree[cur];
            if (!nd.bounds.IntersectP(ray, tMax)) continue;
            if (nd.leaf) {
                for (const index_info& li : nd.triangle_info) {
                    if (model.enable_cull_back_face && !model.back_facing.empty() && model.back_facing[li.mesh_id][li.tri_id]) continue;
                    auto is = model.triangles[li.mesh_id][li.tri_id].BasicIntersect(ray, tMax);
                    if (is.has_value() && is->t < tMax) return true;
                }
            } else {
                for (int i = 0; i < 8; ++i) queue.push(nd.child_id[i]);
            }
        }
        return false;
    }
    int getTreeSize() const { return (int)octtree.size(); }
    struct Stats { int nodes, real_nodes, leaves, empty_leaves, max_leaf; float avg_leaf; int depth; long long refs; };
    Stats GetStats() const {                       // PrintInfo :134-176
        Stats s{(int)octtree.size(), 0, 0, 0, 0, 0, 0, 0};
        std::queue<std::pair<int, int>> q;
        q.push({0, 1});
        long long sum = 0;
        while (!q.empty()) {
            auto [id, d] = q.front();
            q.pop();
            s.real_nodes++;
            s.depth = std::max(s.depth, d);
            if (!octtree[id].leaf) for (int i = 0; i < 8; ++i) q.push({octtree[id].child_id[i], d + 1});
            else {
                sum += (long long)octtree[id].triangle_info.size();
                s.max_leaf = std::max(s.max_leaf, (int)octtree[id].triangle_info.size());
                s.leaves++;
                if (octtree[id].triangle_info.empty()) s.empty_leaves++;
            }
        }
        s.refs = sum;
        s.avg_leaf = sum / (float)s.leaves;
        return s;
    }
};

}  // namespace orc
