// crt/facade.hpp -- the reference's C++ interfaces (GiboDidact/Computational_ray_tracer, render path) on top of the C ABI (crt_b200.h).
//
// A caller of the reference keeps its call sites: the classes below have the reference's names, constructor arguments, virtuals and
// public fields, and forward to libcrt_b200.so.  Header-only, C++17.
//
//   reference (file:line)                                             here
//   Ray, Bounds3, LocalSurfaceInfo          (RayTracer/Shapes.h:33-170)   crt::Ray, crt::Bounds3, crt::LocalSurfaceInfo (all eight fields)
//   class Shape + Sphere / Cylinder / Disk / TriangleSimple (:172-905)    crt::Shape (Bounds / Intersect / IntersectP / Area virtuals) + the four
//   MeshCache::Mesh / Model           (RayTracer/AssetManager.h:20-47)   crt::Mesh, crt::Model (positions, normals, texcoords, tangents, bitangents)
//   TriModel                                 (Shapes.h:1262-1491)        crt::TriModel   (Bounds, ComputeBackFace)
//   Octtree_Model                   (RayTracer/Octtree_Model.h:29-178)   crt::Octtree_Model (CreateOcttree, Traverse(Ray&), getTreeSize, GetNode, PrintInfo)
//   CameraBase::generateRay(vec2, Sampler*), Perspective / Orthographic / Pinhole (RayTracer/Cameras.h:77-359)   same names
//   pbrt::Sampler, IndependentSampler, StratifiedSampler (ThirdParty/pbrv4/samplers.h:25-136)   crt::Sampler + the two
//   pbrt::Filter, BoxFilter, TriangleFilter, GaussianFilter (filters.h:23-163,267-296)          crt::Filter + the three
//   Film                                    (RayTracer/Film.h:6-20)      crt::Film (film_dim, image_res, filter; pixels on the device)
//   Li / evaluate_pixel / thread pool (Applications/RayTracerTestApp.h:218-409), class names of Integrator.h:4-12   crt::Integrator::Render
//
// Vector types: define CRT_FACADE_USE_GLM before including this header to use glm::vec2/vec3/mat4 (what the reference's signatures name;
// glm is not part of the reference repository nor of this one).  Without it, layout-compatible stand-ins (x, y, z members, operator[],
// column-major mat4) are used.
//
// Single-ray calls (Shape::Intersect, Octtree_Model::Traverse(Ray&), CameraBase::generateRay) exist because the reference has them; they cost
// a kernel launch each.  The render path is Integrator::Render; batched forms (TraverseBatch, IntersectBatch) are provided beside them.
// Errors: the reference prints and carries on; here every failing C call throws crt::Error(crt_last_error()).  There is no CPU fallback:
// device methods throw when no CUDA device is present.
#pragma once
#include <array>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../crt_b200.h"

#ifdef CRT_FACADE_USE_GLM
#include <glm/glm.hpp>
#endif

namespace crt {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };
inline void check(int rc) { if (rc != 0) throw Error(crt_last_error()); }

#ifdef CRT_FACADE_USE_GLM
using vec2 = glm::vec2; using vec3 = glm::vec3; using ivec2 = glm::ivec2; using mat4 = glm::mat4;
inline const float* ptr(const mat4& m) { return &m[0][0]; }
inline float* ptr(mat4& m) { return &m[0][0]; }
inline mat4 identity() { return mat4(1.0f); }
#else
struct vec2 { float x = 0, y = 0; vec2() = default; vec2(float x_, float y_) : x(x_), y(y_) {} float& operator[](int i) { return (&x)[i]; } float operator[](int i) const { return (&x)[i]; } };
struct vec3 { float x = 0, y = 0, z = 0; vec3() = default; vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
              float& operator[](int i) { return (&x)[i]; } float operator[](int i) const { return (&x)[i]; } };
struct ivec2 { int x = 0, y = 0; ivec2() = default; ivec2(int x_, int y_) : x(x_), y(y_) {} };
struct mat4 { float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};        // column-major, like glm
              float* operator[](int col) { return m + 4 * col; } const float* operator[](int col) const { return m + 4 * col; } };
inline const float* ptr(const mat4& m) { return m.m; }
inline float* ptr(mat4& m) { return m.m; }
inline mat4 identity() { return mat4(); }
#endif
inline mat4 translation(float x, float y, float z) { mat4 m = identity(); ptr(m)[12] = x; ptr(m)[13] = y; ptr(m)[14] = z; return m; }

struct Ray {                                                       // Shapes.h:33-47
    Ray() = default;
    Ray(vec3 origin, vec3 direction) : o(origin), d(direction) {}
    vec3 o{0, 0, 0};
    vec3 d{0, 0, 1};                                               // normalised
};
struct Bounds3 { vec3 pmin{0, 0, 0}, pmax{0, 0, 0}; };             // Shapes.h:50-139
struct LocalSurfaceInfo {                                          // Shapes.h:144-170
    float tHit = 0;                                                // never assigned by the reference for octree hits (Shapes.h:1034); set by the analytic shapes
    vec3 hitp{0, 0, 0};
    float u = 0, v = 0;
    vec3 du{0, 0, 0}, dv{0, 0, 0};
    vec3 n{0, 0, 0};
    vec3 wo{0, 0, 0};
};

// ---- the process's CUDA context (one per GPU; the reference has no such object, so objects built without one share a default) ---------------
class Context {
public:
    explicit Context(int device = 0) { check(crt_context_create(device, &h_)); }
    ~Context() { crt_context_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    crt_context* handle() const { return h_; }
    void synchronize() { check(crt_context_synchronize(h_)); }
    void set_stream(void* cuda_stream) { check(crt_context_set_stream(h_, cuda_stream)); }
    // pbrt::RGBToSpectrumTable::Init (color.cpp:107-166; RayTracerTestApp.h:138): rebuild the sRGB table on the GPU ...
    void GenerateRgb2Spec() { check(crt_rgb2spec_generate(h_, nullptr, nullptr, nullptr)); }
    // ... or read the reference's own `../rgb2spec/sRGB64binary`
    void LoadRgb2Spec(const std::string& path) {
        std::vector<float> scale(CRT_RGB2SPEC_RES), data(CRT_RGB2SPEC_DATA_FLOATS);
        check(crt_rgb2spec_load_file(path.c_str(), scale.data(), data.data()));
        check(crt_rgb2spec_set(h_, scale.data(), data.data()));
    }
    // Film::pixel_sensor (Film.h:18): the measured-sensor constructor pbrt::PixelSensor(r, g, b, sRGB, illum, ratio) (pixelsensor.h:37-68)
    // from curves sampled at 360..830 nm (471 floats each); UseXYZSensor() = the app's sensor_xyz (RayTracerTestApp.h:149).
    // Returns XYZFromSensorRGB (column-major).  Commit scenes again afterwards.
    std::array<float, 9> SetSensor(const float* r471, const float* g471, const float* b471, const float* illum471, float imagingRatio) {
        std::array<float, 9> m{};
        check(crt_context_set_sensor(h_, r471, g471, b471, illum471, imagingRatio, m.data()));
        return m;
    }
    std::array<float, 9> UseXYZSensor() {
        std::array<float, 9> m{};
        check(crt_context_set_sensor(h_, nullptr, nullptr, nullptr, nullptr, 0.0f, m.data()));
        return m;
    }
    // multi-GPU: one NCCL communicator per context; Film::Reduce sums the per-GPU films onto the root
    static std::array<uint8_t, 128> NcclUniqueId() { std::array<uint8_t, 128> id{}; check(crt_nccl_unique_id(id.data())); return id; }
    void NcclInit(int world, int rank, const std::array<uint8_t, 128>& id) { check(crt_nccl_comm_create(h_, world, rank, id.data())); }
private:
    crt_context* h_ = nullptr;
};
inline Context& DefaultContext() { static Context ctx(0); return ctx; }

// ---- samplers (ThirdParty/pbrv4/samplers.h:25-136) -----------------------------------------------------------------------------------------
// The device's sampler is a value re-seeded per path from (pixel, sample index, dimension, seed); the host classes below reproduce the same
// streams call by call (crt_kat_sampler replays the draws since StartPixelSample), so generateRay(pixel, sampler) behaves like the reference's.
class Sampler {
public:
    virtual ~Sampler() = default;
    virtual int SamplesPerPixel() const = 0;
    virtual void StartPixelSample(ivec2 p, int sampleIndex, int dim = 0) { px_ = p.x; py_ = p.y; index_ = sampleIndex; dim_ = dim; drawn_.clear(); }
    virtual float Get1D() { return draw('1')[0]; }
    virtual vec2 Get2D() { auto v = draw('2'); return vec2(v[0], v[1]); }
    virtual vec2 GetPixel2D() { return Get2D(); }
    // what crt_render_config carries
    int kind = 1, xs = 1, ys = 1, seed = 0;
    bool jitter = true;
protected:
    std::array<float, 2> draw(char what) {
        drawn_.push_back(what);
        std::vector<float> out(2 * drawn_.size());
        check(crt_kat_sampler(kind, xs, ys, jitter ? 1 : 0, seed, px_, py_, index_, dim_, drawn_.c_str(), 0, out.data()));
        size_t at = 0;
        for (size_t i = 0; i + 1 < drawn_.size(); ++i) at += drawn_[i] == '1' ? 1 : 2;
        return {out[at], what == '2' ? out[at + 1] : 0.0f};
    }
    int px_ = 0, py_ = 0, index_ = 0, dim_ = 0;
    std::string drawn_;
};
class IndependentSampler : public Sampler {                       // samplers.h:38-62
public:
    explicit IndependentSampler(int samplesPerPixel, int seed_ = 0) { kind = 0; xs = samplesPerPixel; ys = 1; seed = seed_; }
    int SamplesPerPixel() const override { return xs * ys; }
};
class StratifiedSampler : public Sampler {                        // samplers.h:66-136
public:
    StratifiedSampler(int xPixelSamples, int yPixelSamples, bool jitter_, int seed_ = 0) { kind = 1; xs = xPixelSamples; ys = yPixelSamples; jitter = jitter_; seed = seed_; }
    int SamplesPerPixel() const override { return xs * ys; }
};

// ---- filters (filters.h:23-38,66-163,267-296) ---------------------------------------------------------------------------------------------
struct Filter { int kind = 0; vec2 radius{0.5f, 0.5f}; float sigma = 0.5f; };
struct BoxFilter : Filter { explicit BoxFilter(vec2 r = vec2(0.5f, 0.5f)) { kind = 0; radius = r; } };
struct TriangleFilter : Filter { explicit TriangleFilter(vec2 r = vec2(0.5f, 0.5f)) { kind = 1; radius = r; } };       // deterministic tent, DESIGN.md deviation 1
struct GaussianFilter : Filter { explicit GaussianFilter(vec2 r = vec2(1.5f, 1.5f), float s = 0.5f) { kind = 2; radius = r; sigma = s; } };

// ---- analytic shapes (Shapes.h:172-905) ---------------------------------------------------------------------------------------------------
class Scene;
class Shape {
public:
    Shape(const std::string& _name, mat4 rigidtransform) : name(_name), rigid_(rigidtransform) { update_matrices(); }
    virtual ~Shape() { if (probe_) crt_scene_destroy(probe_); }
    Shape(const Shape&) = delete;
    Shape& operator=(const Shape&) = delete;
    virtual void SetRigidTransform(mat4 rigidtransform) { rigid_ = rigidtransform; update_matrices(); drop_probe(); }
    std::string GetName() const { return name; }
    mat4 GetRenderToObjectMatrix() const { return RenderToObject; }
    mat4 GetObjectToRenderMatrix() const { return ObjectToRender; }
    virtual Bounds3 Bounds() const = 0;
    virtual std::optional<LocalSurfaceInfo> Intersect(const Ray& ray, float tMax = std::numeric_limits<float>::max()) const = 0;
    virtual bool IntersectP(const Ray& ray, float tMax = std::numeric_limits<float>::max()) const = 0;
    virtual float Area() const = 0;
    // how the shape serialises itself into a device scene's shape table (crt_scene_add_shape): 0 Sphere, 1 Cylinder, 2 Disk, 3 TriangleSimple
    virtual int device_kind() const = 0;
    virtual std::array<float, 9> device_params() const = 0;
    const mat4& rigid() const { return rigid_; }
protected:
    // Shape::Intersect / IntersectP for one ray: a private one-shape device scene answers through crt_shape_intersect
    std::optional<LocalSurfaceInfo> device_intersect(const Ray& ray, float tMax) const {
        if (!probe_) {
            check(crt_scene_create(DefaultContext().handle(), &probe_));
            auto p = device_params();
            int id = -1;
            check(crt_scene_add_shape(probe_, device_kind(), ptr(rigid_), p.data(), 0, &id));
            check(crt_scene_commit(probe_));
        }
        const float r[6] = {ray.o[0], ray.o[1], ray.o[2], ray.d[0], ray.d[1], ray.d[2]};
        int32_t found = 0; float t = 0, hp[3] = {0, 0, 0}, n[3] = {0, 0, 0}, uv[2] = {0, 0}, f[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        check(crt_shape_intersect_full(probe_, 0, r, 1, tMax, &found, &t, hp, n, uv, f));
        if (!found) return {};
        LocalSurfaceInfo s;
        s.tHit = t; s.hitp = vec3(hp[0], hp[1], hp[2]); s.n = vec3(n[0], n[1], n[2]); s.u = uv[0]; s.v = uv[1];
        s.du = vec3(f[0], f[1], f[2]); s.dv = vec3(f[3], f[4], f[5]); s.wo = vec3(f[6], f[7], f[8]);
        return s;
    }
    Bounds3 device_bounds() const {
        float b[6]; auto p = device_params();
        check(crt_shape_bounds(device_kind(), ptr(rigid_), p.data(), b));
        Bounds3 r; r.pmin = vec3(b[0], b[1], b[2]); r.pmax = vec3(b[3], b[4], b[5]);
        return r;
    }
    float device_area() const { float a = 0; auto p = device_params(); check(crt_shape_area(device_kind(), p.data(), &a)); return a; }
    void drop_probe() { if (probe_) { crt_scene_destroy(probe_); probe_ = nullptr; } }
    std::string name;
    mat4 RenderToObject, ObjectToRender;
private:
    void update_matrices() { check(crt_shape_matrices(ptr(rigid_), ptr(ObjectToRender), ptr(RenderToObject))); }     // Shapes.h:175-182
    mat4 rigid_;
    mutable crt_scene* probe_ = nullptr;
};
#define CRT_SHAPE_VIRTUALS                                                                                                             \
    Bounds3 Bounds() const override { return device_bounds(); }                                                                        \
    std::optional<LocalSurfaceInfo> Intersect(const Ray& ray, float tMax = std::numeric_limits<float>::max()) const override { return device_intersect(ray, tMax); } \
    bool IntersectP(const Ray& ray, float tMax = std::numeric_limits<float>::max()) const override { return device_intersect(ray, tMax).has_value(); } \
    float Area() const override { return device_area(); }
class Sphere : public Shape {                                                                      // Shapes.h:209-438
public:
    Sphere(const std::string& _name, mat4 rigidtransform, float radius, float _zmin, float _zmax, float _phimax)
        : Shape(_name, rigidtransform), r(radius), zmin(_zmin), zmax(_zmax), phimax(_phimax) {}
    CRT_SHAPE_VIRTUALS
    int device_kind() const override { return 0; }
    std::array<float, 9> device_params() const override { return {r, zmin, zmax, phimax, 0, 0, 0, 0, 0}; }
private:
    float r, zmin, zmax, phimax;
};
class Cylinder : public Shape {                                                                    // Shapes.h:441-629
public:
    Cylinder(const std::string& _name, mat4 rigidtransform, float radius, float _zmin, float _zmax, float _phimax)
        : Shape(_name, rigidtransform), r(radius), zmin(_zmin), zmax(_zmax), phimax(_phimax) {}
    CRT_SHAPE_VIRTUALS
    int device_kind() const override { return 1; }
    std::array<float, 9> device_params() const override { return {r, zmin, zmax, phimax, 0, 0, 0, 0, 0}; }
private:
    float r, zmin, zmax, phimax;
};
class Disk : public Shape {                                                                        // Shapes.h:632-757
public:
    Disk(const std::string& _name, mat4 rigidtransform, float height, float inner_radius, float outer_radius, float _phimax)
        : Shape(_name, rigidtransform), h(height), inner_r(inner_radius), outer_r(outer_radius), phimax(_phimax) {}
    CRT_SHAPE_VIRTUALS
    int device_kind() const override { return 2; }
    std::array<float, 9> device_params() const override { return {h, inner_r, outer_r, phimax, 0, 0, 0, 0, 0}; }
private:
    float h, inner_r, outer_r, phimax;
};
class TriangleSimple : public Shape {                                                              // Shapes.h:760-905
public:
    TriangleSimple(const std::string& _name, mat4 rigidtransform, vec3 _p1, vec3 _p2, vec3 _p3) : Shape(_name, rigidtransform), p1(_p1), p2(_p2), p3(_p3) {}
    CRT_SHAPE_VIRTUALS
    int device_kind() const override { return 3; }
    std::array<float, 9> device_params() const override { return {p1[0], p1[1], p1[2], p2[0], p2[1], p2[2], p3[0], p3[1], p3[2]}; }
    vec3 p1, p2, p3;
};
#undef CRT_SHAPE_VIRTUALS

// ---- triangle models (AssetManager.h:20-47, Shapes.h:1262-1491) -----------------------------------------------------------------------------
struct Mesh {                                                      // AssetManager.h:20-35
    std::vector<float> positions, normals;                         // xyz per vertex; normals may be empty (vertex_available.normals = false)
    std::vector<float> texcoords, tangents, bitangents;            // uv / xyz / xyz per vertex, or empty (the matching vertex_available flag off)
    std::vector<uint32_t> indices;                                 // 3 per triangle
};
struct Model { std::vector<Mesh> meshes; };                        // AssetManager.h:37-47

// MeshCache::LoadMeshFromFile (AssetManager.cpp:8-25,67-190) for Wavefront OBJ
inline Model LoadModelOBJ(const std::string& path) {
    crt_obj* raw = nullptr;
    check(crt_obj_load(path.c_str(), &raw));
    std::unique_ptr<crt_obj, void (*)(crt_obj*)> o(raw, crt_obj_destroy);
    Model model;
    for (int i = 0; i < crt_obj_mesh_count(o.get()); ++i) {
        uint32_t nv = 0, nt = 0;
        check(crt_obj_mesh_info(o.get(), i, &nv, &nt, nullptr, 0));
        Mesh m;
        m.positions.resize(3 * (size_t)nv); m.normals.resize(3 * (size_t)nv); m.indices.resize(3 * (size_t)nt);
        check(crt_obj_mesh_copy(o.get(), i, m.positions.data(), m.normals.data(), m.indices.data()));
        std::vector<float> uv(2 * (size_t)nv), tan(3 * (size_t)nv), bitan(3 * (size_t)nv);
        int available = 0;
        check(crt_obj_mesh_attributes(o.get(), i, uv.data(), tan.data(), bitan.data(), &available));
        if (available) { m.texcoords = std::move(uv); m.tangents = std::move(tan); m.bitangents = std::move(bitan); }
        model.meshes.push_back(std::move(m));
    }
    return model;
}

// TriModel: a Model + rigid transform + culling table (Shapes.h:1282-1397)
class TriModel {
public:
    TriModel(const Model& model, const mat4& rigidtransform, bool cull_back_face, bool precomputed_worldtransform)
        : model_(model), cull_(cull_back_face), precomputed_(precomputed_worldtransform) {
        mat4 r2o;
        check(crt_shape_matrices(ptr(rigidtransform), ptr(o2r_), ptr(r2o)));
        rebuild_descs();
    }
    TriModel(const TriModel& other) : model_(other.model_), o2r_(other.o2r_), cull_(other.cull_), precomputed_(other.precomputed_), back_facing_(other.back_facing_) { rebuild_descs(); }
    TriModel& operator=(const TriModel& other) {
        if (this != &other) { model_ = other.model_; o2r_ = other.o2r_; cull_ = other.cull_; precomputed_ = other.precomputed_; back_facing_ = other.back_facing_; rebuild_descs(); }
        return *this;
    }
    // Shapes.h:1339-1380
    void ComputeBackFace(const vec3& look_direction, bool enable) {
        cull_ = enable;
        back_facing_.clear();
        const float look[3] = {look_direction[0], look_direction[1], look_direction[2]};
        for (size_t m = 0; m < descs_.size(); ++m) {
            std::vector<uint8_t> bits(descs_[m].n_triangles);
            check(crt_model_compute_backface(&descs_[m], look, ptr(o2r_), precomputed_ ? 1 : 0, bits.data()));
            back_facing_.push_back(std::move(bits));
        }
    }
    Bounds3 Bounds() const {                                        // Shapes.h:1390-1397
        float b[6];
        check(crt_model_bounds(descs_.data(), (uint32_t)descs_.size(), ptr(o2r_), precomputed_ ? 1 : 0, b));
        Bounds3 r; r.pmin = vec3(b[0], b[1], b[2]); r.pmax = vec3(b[3], b[4], b[5]);
        return r;
    }
    const std::vector<crt_mesh_desc>& descs() const { return descs_; }
    const mat4& ObjectToRender() const { return o2r_; }
    bool precomputed() const { return precomputed_; }
    bool culling() const { return cull_ && !back_facing_.empty(); }
    const std::vector<std::vector<uint8_t>>& back_facing() const { return back_facing_; }
private:
    void rebuild_descs() {                                          // descs_ point into this object's own model_
        descs_.clear();
        for (const Mesh& m : model_.meshes) {
            crt_mesh_desc d{};
            d.positions = m.positions.data(); d.normals = m.normals.empty() ? nullptr : m.normals.data();
            d.n_vertices = (uint32_t)(m.positions.size() / 3); d.indices = m.indices.data(); d.n_triangles = (uint32_t)(m.indices.size() / 3);
            d.texcoords = m.texcoords.empty() ? nullptr : m.texcoords.data();
            d.tangents = m.tangents.empty() ? nullptr : m.tangents.data();
            d.bitangents = m.bitangents.empty() ? nullptr : m.bitangents.data();
            descs_.push_back(d);
        }
    }
    Model model_;
    std::vector<crt_mesh_desc> descs_;
    mat4 o2r_;
    bool cull_, precomputed_;
    std::vector<std::vector<uint8_t>> back_facing_;
};

// Octtree_Model.h:29-178
class Octtree_Model {
public:
    explicit Octtree_Model(TriModel& model) : model(model) {}
    ~Octtree_Model() { if (probe_) crt_scene_destroy(probe_); crt_octree_destroy(h_); }
    Octtree_Model(const Octtree_Model&) = delete;
    Octtree_Model& operator=(const Octtree_Model&) = delete;
    void CreateOcttree() {                                          // :33-63 -- the reference's insertion algorithm on the host
        reset();
        check(crt_octree_build(model.descs().data(), (uint32_t)model.descs().size(), ptr(model.ObjectToRender()), model.precomputed() ? 1 : 0, &h_));
    }
    // the same tree built on the GPU (level-synchronous, csrc/crt_build.cuh); node ids are breadth-first
    void CreateOcttree(Context& ctx) {
        reset();
        check(crt_octree_build_gpu(ctx.handle(), model.descs().data(), (uint32_t)model.descs().size(), ptr(model.ObjectToRender()), model.precomputed() ? 1 : 0, &h_));
    }
    int getTreeSize() const { return crt_octree_node_count(h_); }  // :129
    struct node { Bounds3 bounds; bool leaf = true; std::array<int32_t, 8> child_id{}; std::vector<std::array<int32_t, 2>> triangle_info; };
    node GetNode(int i) const {                                     // :178
        node n; int32_t leaf = 0, cnt = 0; float b[6];
        check(crt_octree_get_node(h_, i, b, &leaf, n.child_id.data(), nullptr, 0, &cnt));
        n.triangle_info.resize(cnt);
        if (cnt) check(crt_octree_get_node(h_, i, b, &leaf, n.child_id.data(), &n.triangle_info[0][0], cnt, &cnt));
        n.bounds.pmin = vec3(b[0], b[1], b[2]); n.bounds.pmax = vec3(b[3], b[4], b[5]);
        n.leaf = leaf != 0;
        return n;
    }
    void PrintInfo() const {                                        // :134-176
        crt_octree_stats s;
        check(crt_octree_get_stats(h_, &s));
        std::printf("octree: %d nodes, %d leaves (%d empty), avg %.2f / max %d triangles per leaf, depth %d\n", s.nodes, s.leaves, s.empty_leaves, s.avg_leaf, s.max_leaf, s.depth);
    }
    // Traverse(Ray&) (:66-127): closest hit over the octree and the full LocalSurfaceInfo of Triangle::CalculateLocalSurface, on the GPU.
    // The flattened tree is uploaded to the default context on first use ("Flatten() + Upload()").
    std::optional<LocalSurfaceInfo> Traverse(Ray& ray) const {
        std::vector<std::optional<LocalSurfaceInfo>> r = TraverseBatch(&ray, 1);
        return r[0];
    }
    std::vector<std::optional<LocalSurfaceInfo>> TraverseBatch(const Ray* rays, int n) const {
        upload();
        std::vector<float> r6(6 * (size_t)n), info(17 * (size_t)n);
        std::vector<int32_t> found(n);
        for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) { r6[6 * i + k] = rays[i].o[k]; r6[6 * i + 3 + k] = rays[i].d[k]; }
        check(crt_traverse_local_surface(probe_, r6.data(), n, 3, found.data(), info.data()));
        std::vector<std::optional<LocalSurfaceInfo>> out(n);
        for (int i = 0; i < n; ++i) {
            if (!found[i]) continue;
            const float* w = &info[17 * (size_t)i];
            LocalSurfaceInfo s;
            s.hitp = vec3(w[0], w[1], w[2]); s.u = w[3]; s.v = w[4]; s.du = vec3(w[5], w[6], w[7]); s.dv = vec3(w[8], w[9], w[10]);
            s.n = vec3(w[11], w[12], w[13]); s.wo = vec3(w[14], w[15], w[16]);
            out[i] = s;
        }
        return out;
    }
    const crt_octree* handle() const { return h_; }
    TriModel& model;                                                // :416
private:
    void reset() { if (probe_) { crt_scene_destroy(probe_); probe_ = nullptr; } crt_octree_destroy(h_); h_ = nullptr; }
    void upload() const {
        if (probe_) return;
        if (!h_) throw Error("Octtree_Model::Traverse: call CreateOcttree() first");
        check(crt_scene_create(DefaultContext().handle(), &probe_));
        std::vector<const uint8_t*> cull;
        if (model.culling()) for (auto& b : model.back_facing()) cull.push_back(b.data());
        check(crt_scene_set_model(probe_, model.descs().data(), (uint32_t)model.descs().size(), ptr(model.ObjectToRender()), model.precomputed() ? 1 : 0,
                                  cull.empty() ? nullptr : cull.data(), h_, nullptr));
        check(crt_scene_commit(probe_));
    }
    crt_octree* h_ = nullptr;
    mutable crt_scene* probe_ = nullptr;
};

// ---- device-resident scene: the flattened octree + triangles + the shape table (+ materials for the path integrator) ------------------------
class Scene {
public:
    explicit Scene(Context& ctx = DefaultContext()) : ctx_(ctx) { check(crt_scene_create(ctx.handle(), &h_)); }
    ~Scene() { crt_scene_destroy(h_); }
    Scene(const Scene&) = delete;
    Scene& operator=(const Scene&) = delete;
    void SetModel(const Octtree_Model& oct, const std::vector<int32_t>& mesh_materials = {}) {
        const TriModel& m = oct.model;
        std::vector<const uint8_t*> cull;
        if (m.culling()) for (auto& b : m.back_facing()) cull.push_back(b.data());
        check(crt_scene_set_model(h_, m.descs().data(), (uint32_t)m.descs().size(), ptr(m.ObjectToRender()), m.precomputed() ? 1 : 0,
                                  cull.empty() ? nullptr : cull.data(), oct.handle(), mesh_materials.empty() ? nullptr : mesh_materials.data()));
    }
    // any Shape serialises itself into the device shape table
    int Add(const Shape& shape, int material = 0) {
        auto p = shape.device_params(); int id = -1;
        check(crt_scene_add_shape(h_, shape.device_kind(), ptr(shape.rigid()), p.data(), material, &id));
        return id;
    }
    int AddConstantSpectrum(float c) { int id; check(crt_scene_add_spectrum(h_, 0, c, nullptr, 0, nullptr, 0, &id)); return id; }
    int AddPiecewiseLinearSpectrum(const std::vector<float>& interleaved, bool normalize) {                        // spectrum.cpp:134-165
        int id; check(crt_scene_add_spectrum(h_, 1, 0, interleaved.data(), (int)interleaved.size(), nullptr, normalize, &id)); return id;
    }
    int AddNamedSpectrum(const std::string& name) { int id; check(crt_scene_add_spectrum(h_, 2, 0, nullptr, 0, name.c_str(), 0, &id)); return id; }
    int AddStdIlluminant(int which) { int id; check(crt_scene_add_spectrum(h_, 4, 0, nullptr, which, nullptr, 0, &id)); return id; }
    int AddMaterial(int type, int refl, int eta = -1, int k = -1, int emit = -1, float emit_scale = 0, bool two_sided = false, bool eta_constant = true) {
        int id; check(crt_scene_add_material(h_, type, refl, eta, k, emit, emit_scale, two_sided, eta_constant, &id)); return id;
    }
    // Lights.h:5-8: "points light: position, color, and r^2 falloff" / "sunlight: direction, color"
    int AddPointLight(const vec3& position, int spectrum, float intensity_scale) {
        const float v[3] = {position[0], position[1], position[2]}; int id; check(crt_scene_add_light(h_, 0, v, spectrum, intensity_scale, &id)); return id;
    }
    int AddSunLight(const vec3& direction_to_light, int spectrum, float irradiance_scale) {
        const float v[3] = {direction_to_light[0], direction_to_light[1], direction_to_light[2]}; int id; check(crt_scene_add_light(h_, 1, v, spectrum, irradiance_scale, &id)); return id;
    }
    void Commit() { check(crt_scene_commit(h_)); }
    crt_scene* handle() const { return h_; }
    Context& context() const { return ctx_; }
private:
    Context& ctx_;
    crt_scene* h_ = nullptr;
};

// ---- cameras (Cameras.h:77-359).  The matrices are computed by the library's host code (bit-identical to the reference's). -------------------
class CameraBase {
public:
    virtual ~CameraBase() = default;
    // Cameras.h:179: one camera ray through a film position; a thin lens consumes sampler->Get2D()
    virtual Ray generateRay(vec2 pixel, Sampler* sampler) {
        float xy[2] = {pixel[0], pixel[1]}, u[2] = {0, 0}, r[6];
        const bool lens = kind == 0 && lensRadius > 0 && sampler;
        if (lens) { vec2 s = sampler->Get2D(); u[0] = s[0]; u[1] = s[1]; }
        check(crt_camera_generate_rays(kind, ptr(M_RastertoCamera), ptr(M_CameratoWorld), lensRadius, focalDistance, xy, lens ? u : nullptr, 1, 0, r));
        return Ray(vec3(r[0], r[1], r[2]), vec3(r[3], r[4], r[5]));
    }
    void SetlensRadius(float r) { lensRadius = r; }
    void SetfocalDistance(float d) { focalDistance = d; }
    void setWorldPos(vec3 p) { world_pos = p; recompute(); }
    void setLookDirection(vec3 d) { look_direction = d; recompute(); }
    mat4 M_RastertoCamera, M_CameratoWorld;
    float lensRadius = 0, focalDistance = 0;
    int kind = 0;                                                   // 0 Perspective, 1 Orthographic, 2 Pinhole (crt_render_config.camera_kind)
protected:
    void recompute() {
        const float pos[3] = {world_pos[0], world_pos[1], world_pos[2]}, look[3] = {look_direction[0], look_direction[1], look_direction[2]};
        const float right[3] = {1, 0, 0}, up[3] = {worldup_direction[0], worldup_direction[1], worldup_direction[2]};
        check(crt_camera_matrices(kind, near_, far_, sensor_w_, sensor_h_, fov_, pos, look, right, up, res_x_, res_y_, ptr(M_RastertoCamera), ptr(M_CameratoWorld)));
    }
    vec3 world_pos{0, 0, 0}, look_direction{0, 0, 1}, worldup_direction{0, 1, 0};
    float near_ = 1, far_ = 1000, sensor_w_ = 0, sensor_h_ = 0, fov_ = 45, res_x_ = 1, res_y_ = 1;
};
class PerspectiveCamera : public CameraBase {                                                // Cameras.h:248-311
public:
    PerspectiveCamera(float nearp, float farp, float fov_deg, vec3 pos, vec3 look, vec3 worldup, float res_x, float res_y, float lens_radius = 0, float focal_distance = 0) {
        kind = 0; near_ = nearp; far_ = farp; fov_ = fov_deg; world_pos = pos; look_direction = look; worldup_direction = worldup; res_x_ = res_x; res_y_ = res_y;
        lensRadius = lens_radius; focalDistance = focal_distance;
        recompute();
    }
    void ChangeFOV(float fov_deg) { fov_ = fov_deg; recompute(); }                            // Cameras.h:299-301
};
class OrthographicCamera : public CameraBase {                                               // Cameras.h:213-245
public:
    OrthographicCamera(float nearp, float farp, float sensor_w, float sensor_h, vec3 pos, vec3 look, vec3 worldup, float res_x, float res_y) {
        kind = 1; near_ = nearp; far_ = farp; sensor_w_ = sensor_w; sensor_h_ = sensor_h; world_pos = pos; look_direction = look; worldup_direction = worldup;
        res_x_ = res_x; res_y_ = res_y;
        recompute();
    }
};
class PinholeCamera : public CameraBase {                                                    // Cameras.h:313-359
public:
    PinholeCamera(float /*hole_radius*/, vec3 box_dimensions, vec3 pos, vec3 look, vec3 worldup, float res_x, float res_y) {
        kind = 2; near_ = 0; far_ = 0; sensor_w_ = box_dimensions[0]; sensor_h_ = box_dimensions[1]; world_pos = pos; look_direction = look; worldup_direction = worldup;
        res_x_ = res_x; res_y_ = res_y;
        focalDistance = box_dimensions[2];                          // M_RastertoCamera carries M_RastertoScreen for this camera
        recompute();
    }
};

// ---- Film.h:6-20: pixels = (rgbsum, weightsum); the device copy is authoritative during a render ------------------------------------------
class Film {
public:
    Film(int width, int height, Filter* filter_ = nullptr, Context& ctx = DefaultContext()) : film_dim(width, height), image_res(width, height), filter(filter_) {
        check(crt_film_create(ctx.handle(), width, height, &f_));
    }
    ~Film() { crt_film_destroy(f_); }
    Film(const Film&) = delete;
    Film& operator=(const Film&) = delete;
    void Clear() { check(crt_film_clear(f_)); }                                              // "restart" button, RayTracerTestApp.h:490-495
    std::vector<float> Pixels() const { std::vector<float> p((size_t)image_res.x * image_res.y * 4); check(crt_film_download(f_, p.data())); return p; }   // Film::pixels
    void Restore(const std::vector<float>& p) { check(crt_film_upload(f_, p.data())); }      // checkpoint / resume
    std::vector<uint8_t> ResolveRGB8() const { std::vector<uint8_t> o((size_t)image_res.x * image_res.y * 3); check(crt_film_resolve(f_, o.data(), nullptr)); return o; }   // :425-452
    void Reduce(int root = 0) { check(crt_film_reduce(f_, root)); }                          // multi-GPU: ncclReduce(sum) onto the root
    int width() const { return image_res.x; }
    int height() const { return image_res.y; }
    crt_film* handle() const { return f_; }
    ivec2 film_dim, image_res;
    Filter* filter;                                                                           // Film.h:19 (nullptr = BoxFilter(0.5))
private:
    crt_film* f_ = nullptr;
};

// ---- the integrator entry (class names from Integrator.h:4-12): SimplePathIntegrator ~ mode 0 (the reference's Li), PathIntegrator ~ mode 1 ---
struct Integrator {
    int mode = 0, max_depth = 5, rr_depth = 0;
    float ray_eps = 1e-2f, shadow_eps = 1e-3f;
    vec3 albedo{0.5f, 0.5f, 0.5f};                                                            // the app's `colors` (RayTracerTestApp.h:207)
    int rank = 0, world = 1, partition = 1, tile_w = 32, tile_h = 32;
    int trace_mode = 3;
    int shade_mode = 0;                                                                        // 0 automatic, 1 fused, 2 staged per material type (identical films)
    int light_strategy = 0;                                                                    // 0 power CDF, 1 "1 sample from each light source" (Shading.h:4)

    crt_render_config Config(const Film& film, const CameraBase& cam, const Sampler& s, int spp_begin, int spp_end) const {
        crt_render_config c{};
        c.width = film.width(); c.height = film.height();
        std::memcpy(c.raster_to_camera, ptr(cam.M_RastertoCamera), 64); std::memcpy(c.camera_to_world, ptr(cam.M_CameratoWorld), 64);
        c.lens_radius = cam.lensRadius; c.focal_distance = cam.focalDistance; c.camera_kind = cam.kind;
        c.sampler_kind = s.kind; c.xs = s.xs; c.ys = s.ys; c.jitter = s.jitter ? 1 : 0; c.seed = s.seed;
        const Filter box;
        const Filter& f = film.filter ? *film.filter : box;
        c.filter_kind = f.kind; c.filter_rx = f.radius[0]; c.filter_ry = f.radius[1]; c.filter_sigma = f.sigma;
        c.mode = mode; c.max_depth = max_depth; c.rr_depth = rr_depth; c.ray_eps = ray_eps; c.shadow_eps = shadow_eps;
        for (int i = 0; i < 3; ++i) c.albedo[i] = albedo[i];
        c.spp_begin = spp_begin; c.spp_end = spp_end; c.rank = rank; c.world = world; c.partition = partition; c.tile_w = tile_w; c.tile_h = tile_h;
        c.trace_mode = trace_mode; c.light_strategy = light_strategy; c.shade_mode = shade_mode;
        return c;
    }
    // evaluate_pixel for every pixel and sample index in [spp_begin, spp_end) (RayTracerTestApp.h:287-409)
    crt_render_stats Render(Scene& scene, Film& film, const CameraBase& cam, const Sampler& sampler, int spp_begin, int spp_end) const {
        crt_render_config c = Config(film, cam, sampler, spp_begin, spp_end);
        crt_render_stats st{};
        check(crt_render(scene.handle(), film.handle(), &c, &st));
        return st;
    }
};

}  // namespace crt
