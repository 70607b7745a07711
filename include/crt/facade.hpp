// crt/facade.hpp -- C++ facade over the C ABI (crt_b200.h) with the reference's class shapes, so that a caller of
// GiboDidact/Computational_ray_tracer's render path can switch by changing includes.  Header-only, C++17, no glm:
// vectors/matrices are plain float arrays (column-major like glm, RayTracer/Shapes.h:175-182).
//
//   reference                                             facade
//   MeshCache::Mesh / Model (AssetManager.h:20-47)        crt::Mesh, crt::Model
//   TriModel (Shapes.h:1262-1491)                         crt::TriModel        (Bounds, ComputeBackFace)
//   Octtree_Model (Octtree_Model.h:29-178)                crt::Octtree_Model   (CreateOcttree, Traverse, getTreeSize, GetNode, PrintInfo)
//   PerspectiveCamera / OrthographicCamera (Cameras.h)    crt::PerspectiveCamera, crt::OrthographicCamera
//   pbrt::StratifiedSampler / IndependentSampler          crt::SamplerDesc
//   pbrt::BoxFilter / TriangleFilter                      crt::FilterDesc
//   Film (Film.h:6-20) + resolve (RayTracerTestApp.h:425) crt::Film
//   Li / evaluate_pixel / thread pool (RayTracerTestApp)  crt::Integrator::Render
//
// Errors: the reference prints and carries on; here every failing C call throws crt::Error(crt_last_error()).
#pragma once
#include <array>
#include <cstdint>
#include <cstdio>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../crt_b200.h"

namespace crt {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };
inline void check(int rc) { if (rc != 0) throw Error(crt_last_error()); }

using vec3 = std::array<float, 3>;
using mat4 = std::array<float, 16>;        // column-major
inline mat4 identity() { return {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; }

struct Ray { vec3 o{0, 0, 0}, d{0, 0, 1}; };                      // Shapes.h:33-47
struct LocalSurfaceInfo { float tHit = 0; vec3 n{0, 0, 0}; int mesh_id = -1, tri_id = -1; vec3 bary{0, 0, 0}; };   // Shapes.h:144-170 (fields Li reads)

struct Mesh {                                                      // AssetManager.h:20-35
    std::vector<float> positions, normals;                         // xyz per vertex; normals may be empty
    std::vector<uint32_t> indices;                                 // 3 per triangle
};
struct Model { std::vector<Mesh> meshes; };                        // AssetManager.h:37-47

// MeshCache::LoadMeshFromFile (AssetManager.cpp:8-25) for Wavefront OBJ
inline Model LoadModelOBJ(const std::string& path) {
    crt_obj* o = nullptr;
    check(crt_obj_load(path.c_str(), &o));
    Model model;
    for (int i = 0; i < crt_obj_mesh_count(o); ++i) {
        uint32_t nv = 0, nt = 0;
        check(crt_obj_mesh_info(o, i, &nv, &nt, nullptr, 0));
        Mesh m;
        m.positions.resize(3 * (size_t)nv); m.normals.resize(3 * (size_t)nv); m.indices.resize(3 * (size_t)nt);
        check(crt_obj_mesh_copy(o, i, m.positions.data(), m.normals.data(), m.indices.data()));
        model.meshes.push_back(std::move(m));
    }
    crt_obj_destroy(o);
    return model;
}

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
    void LoadRgb2Spec(const std::string& path) {
        std::vector<float> scale(CRT_RGB2SPEC_RES), data(CRT_RGB2SPEC_DATA_FLOATS);
        check(crt_rgb2spec_load_file(path.c_str(), scale.data(), data.data()));
        check(crt_rgb2spec_set(h_, scale.data(), data.data()));
    }
private:
    crt_context* h_ = nullptr;
};

// TriModel: a Model + rigid transform + culling table (Shapes.h:1282-1397)
class TriModel {
public:
    TriModel(const Model& model, const mat4& rigidtransform, bool cull_back_face, bool precomputed_worldtransform)
        : model_(model), cull_(cull_back_face), precomputed_(precomputed_worldtransform) {
        mat4 r2o;
        check(crt_shape_matrices(rigidtransform.data(), o2r_.data(), r2o.data()));
        for (const Mesh& m : model_.meshes) {
            crt_mesh_desc d;
            d.positions = m.positions.data(); d.normals = m.normals.empty() ? nullptr : m.normals.data();
            d.n_vertices = (uint32_t)(m.positions.size() / 3); d.indices = m.indices.data(); d.n_triangles = (uint32_t)(m.indices.size() / 3);
            descs_.push_back(d);
        }
    }
    // Shapes.h:1339-1380
    void ComputeBackFace(const vec3& look_direction, bool enable) {
        cull_ = enable;
        back_facing_.clear();
        for (size_t m = 0; m < descs_.size(); ++m) {
            std::vector<uint8_t> bits(descs_[m].n_triangles);
            check(crt_model_compute_backface(&descs_[m], look_direction.data(), o2r_.data(), precomputed_ ? 1 : 0, bits.data()));
            back_facing_.push_back(std::move(bits));
        }
    }
    std::array<float, 6> Bounds() const {                           // Shapes.h:1390-1397
        std::array<float, 6> b{};
        check(crt_model_bounds(descs_.data(), (uint32_t)descs_.size(), o2r_.data(), precomputed_ ? 1 : 0, b.data()));
        return b;
    }
    const std::vector<crt_mesh_desc>& descs() const { return descs_; }
    const mat4& ObjectToRender() const { return o2r_; }
    bool precomputed() const { return precomputed_; }
    bool culling() const { return cull_ && !back_facing_.empty(); }
    const std::vector<std::vector<uint8_t>>& back_facing() const { return back_facing_; }
private:
    Model model_;
    std::vector<crt_mesh_desc> descs_;
    mat4 o2r_{};
    bool cull_, precomputed_;
    std::vector<std::vector<uint8_t>> back_facing_;
};

class Scene;

// Octtree_Model.h:29-178
class Octtree_Model {
public:
    explicit Octtree_Model(TriModel& model) : model_(model) {}
    ~Octtree_Model() { crt_octree_destroy(h_); }
    Octtree_Model(const Octtree_Model&) = delete;
    void CreateOcttree() {                                          // :33-63
        crt_octree_destroy(h_); h_ = nullptr;
        check(crt_octree_build(model_.descs().data(), (uint32_t)model_.descs().size(), model_.ObjectToRender().data(), model_.precomputed() ? 1 : 0, &h_));
    }
    // the same tree built on the GPU (level-synchronous, csrc/crt_build.cuh); node ids are breadth-first
    void CreateOcttree(Context& ctx) {
        crt_octree_destroy(h_); h_ = nullptr;
        check(crt_octree_build_gpu(ctx.handle(), model_.descs().data(), (uint32_t)model_.descs().size(), model_.ObjectToRender().data(), model_.precomputed() ? 1 : 0, &h_));
    }
    int getTreeSize() const { return crt_octree_node_count(h_); }  // :129
    struct node { std::array<float, 6> bounds; bool leaf; std::array<int32_t, 8> child_id; std::vector<std::array<int32_t, 2>> triangle_info; };
    node GetNode(int i) const {                                     // :178
        node n; int32_t leaf = 0, cnt = 0;
        check(crt_octree_get_node(h_, i, n.bounds.data(), &leaf, n.child_id.data(), nullptr, 0, &cnt));
        n.triangle_info.resize(cnt);
        if (cnt) check(crt_octree_get_node(h_, i, n.bounds.data(), &leaf, n.child_id.data(), &n.triangle_info[0][0], cnt, &cnt));
        n.leaf = leaf != 0;
        return n;
    }
    void PrintInfo() const {                                        // :134-176
        crt_octree_stats s;
        check(crt_octree_get_stats(h_, &s));
        std::printf("octree: %d nodes, %d leaves (%d empty), avg %.2f / max %d triangles per leaf, depth %d\n", s.nodes, s.leaves, s.empty_leaves, s.avg_leaf, s.max_leaf, s.depth);
    }
    // Traverse(Ray&) (:66-127) for one ray or a batch, on the GPU; defined after Scene
    std::optional<LocalSurfaceInfo> Traverse(Scene& scene, const Ray& ray) const;
    const crt_octree* handle() const { return h_; }
    TriModel& model() const { return model_; }
private:
    TriModel& model_;
    crt_octree* h_ = nullptr;
};

// Device-resident scene: the flattened octree + triangles (+ shapes/materials for the path integrator)
class Scene {
public:
    explicit Scene(Context& ctx) : ctx_(ctx) { check(crt_scene_create(ctx.handle(), &h_)); }
    ~Scene() { crt_scene_destroy(h_); }
    Scene(const Scene&) = delete;
    void SetModel(const Octtree_Model& oct, const std::vector<int32_t>& mesh_materials = {}) {
        const TriModel& m = oct.model();
        std::vector<const uint8_t*> cull;
        if (m.culling()) for (auto& b : m.back_facing()) cull.push_back(b.data());
        check(crt_scene_set_model(h_, m.descs().data(), (uint32_t)m.descs().size(), m.ObjectToRender().data(), m.precomputed() ? 1 : 0,
                                  cull.empty() ? nullptr : cull.data(), oct.handle(), mesh_materials.empty() ? nullptr : mesh_materials.data()));
    }
    int AddSphere(const mat4& rigid, float radius, float zmin, float zmax, float phimax_deg, int material = 0) {   // Shapes.h:209-231
        float p[9] = {radius, zmin, zmax, phimax_deg}; int id = -1;
        check(crt_scene_add_shape(h_, 0, rigid.data(), p, material, &id)); return id;
    }
    int AddCylinder(const mat4& rigid, float radius, float zmin, float zmax, float phimax_deg, int material = 0) { // Shapes.h:441-466
        float p[9] = {radius, zmin, zmax, phimax_deg}; int id = -1;
        check(crt_scene_add_shape(h_, 1, rigid.data(), p, material, &id)); return id;
    }
    int AddDisk(const mat4& rigid, float height, float inner_r, float outer_r, float phimax_deg, int material = 0) { // Shapes.h:632-655
        float p[9] = {height, inner_r, outer_r, phimax_deg}; int id = -1;
        check(crt_scene_add_shape(h_, 2, rigid.data(), p, material, &id)); return id;
    }
    int AddTriangleSimple(const mat4& rigid, const vec3& p1, const vec3& p2, const vec3& p3, int material = 0) {   // Shapes.h:776-795
        float p[9] = {p1[0], p1[1], p1[2], p2[0], p2[1], p2[2], p3[0], p3[1], p3[2]}; int id = -1;
        check(crt_scene_add_shape(h_, 3, rigid.data(), p, material, &id)); return id;
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
    void Commit() { check(crt_scene_commit(h_)); }
    crt_scene* handle() const { return h_; }
private:
    Context& ctx_;
    crt_scene* h_ = nullptr;
};

inline std::optional<LocalSurfaceInfo> Octtree_Model::Traverse(Scene& scene, const Ray& ray) const {
    float r[6] = {ray.o[0], ray.o[1], ray.o[2], ray.d[0], ray.d[1], ray.d[2]};
    int32_t mesh = -1, tri = -1, found = 0; float t = 0, b[3] = {0, 0, 0}, n[3] = {0, 0, 0};
    check(crt_trace_closest(scene.handle(), r, 1, 0, &mesh, &tri, &t, b));
    if (tri < 0) return {};
    check(crt_traverse_surface(scene.handle(), r, 1, &found, n));
    LocalSurfaceInfo s; s.tHit = t; s.n = {n[0], n[1], n[2]}; s.mesh_id = mesh; s.tri_id = tri; s.bary = {b[0], b[1], b[2]};
    return s;
}

// Cameras.h:77-311.  The matrices are computed by the library's host code (same evaluation order as the oracle).
struct CameraBase {
    mat4 M_RastertoCamera{}, M_CameratoWorld{};
    float lensRadius = 0, focalDistance = 0;
    int kind = 0;
    void SetlensRadius(float r) { lensRadius = r; }
    void SetfocalDistance(float d) { focalDistance = d; }
};
struct PerspectiveCamera : CameraBase {                                                     // Cameras.h:248-311
    PerspectiveCamera(float near_, float far_, float fov_deg, const vec3& pos, const vec3& look, const vec3& worldup, float res_x, float res_y,
                      float lens_radius = 0, float focal_distance = 0) {
        vec3 right{1, 0, 0};
        check(crt_camera_matrices(0, near_, far_, 0, 0, fov_deg, pos.data(), look.data(), right.data(), worldup.data(), res_x, res_y, M_RastertoCamera.data(), M_CameratoWorld.data()));
        lensRadius = lens_radius; focalDistance = focal_distance; kind = 0;
    }
};
struct OrthographicCamera : CameraBase {                                                    // Cameras.h:213-245
    OrthographicCamera(float near_, float far_, float sensor_w, float sensor_h, const vec3& pos, const vec3& look, const vec3& worldup, float res_x, float res_y) {
        vec3 right{1, 0, 0};
        check(crt_camera_matrices(1, near_, far_, sensor_w, sensor_h, 0, pos.data(), look.data(), right.data(), worldup.data(), res_x, res_y, M_RastertoCamera.data(), M_CameratoWorld.data()));
        kind = 1;
    }
};

struct PinholeCamera : CameraBase {                                                         // Cameras.h:313-359
    PinholeCamera(float /*hole_radius*/, const vec3& box_dimensions, const vec3& pos, const vec3& look, const vec3& worldup, float res_x, float res_y) {
        vec3 right{1, 0, 0};
        check(crt_camera_matrices(2, 0, 0, box_dimensions[0], box_dimensions[1], 0, pos.data(), look.data(), right.data(), worldup.data(), res_x, res_y, M_RastertoCamera.data(), M_CameratoWorld.data()));
        focalDistance = box_dimensions[2]; kind = 2;        // M_RastertoCamera carries M_RastertoScreen for this camera
    }
};

struct SamplerDesc { int kind = 1, xs = 4, ys = 4; bool jitter = true; int seed = 0; };      // samplers.h:38-136
struct FilterDesc { int kind = 0; float rx = 0.5f, ry = 0.5f; float sigma = 0.5f; };         // 0 Box, 1 Triangle, 2 Gaussian(sigma); filters.h:66-163,267-296

// Film.h:6-20: pixels = (rgbsum, weightsum); the device copy is authoritative during a render.
class Film {
public:
    Film(Context& ctx, int width, int height) : w_(width), h_(height) { check(crt_film_create(ctx.handle(), width, height, &f_)); }
    ~Film() { crt_film_destroy(f_); }
    Film(const Film&) = delete;
    void Clear() { check(crt_film_clear(f_)); }                                              // "restart" button, RayTracerTestApp.h:490-495
    std::vector<float> Pixels() const { std::vector<float> p((size_t)w_ * h_ * 4); check(crt_film_download(f_, p.data())); return p; }
    void Restore(const std::vector<float>& p) { check(crt_film_upload(f_, p.data())); }      // checkpoint / resume
    std::vector<uint8_t> ResolveRGB8() const { std::vector<uint8_t> o((size_t)w_ * h_ * 3); check(crt_film_resolve(f_, o.data(), nullptr)); return o; }   // :425-452
    int width() const { return w_; }
    int height() const { return h_; }
    crt_film* handle() const { return f_; }
private:
    int w_, h_;
    crt_film* f_ = nullptr;
};

// The integrator entry (class names from Integrator.h:4-12): SimplePathIntegrator ~ mode 0 (the reference's Li),
// PathIntegrator ~ mode 1 (NEE + BSDF sampling).
struct Integrator {
    int mode = 0, max_depth = 5, rr_depth = 0;
    float ray_eps = 1e-2f, shadow_eps = 1e-3f;
    vec3 albedo{0.5f, 0.5f, 0.5f};
    int rank = 0, world = 1, partition = 1, tile_w = 32, tile_h = 32;

    crt_render_config Config(const Film& film, const CameraBase& cam, const SamplerDesc& s, const FilterDesc& f, int spp_begin, int spp_end) const {
        crt_render_config c{};
        c.width = film.width(); c.height = film.height();
        for (int i = 0; i < 16; ++i) { c.raster_to_camera[i] = cam.M_RastertoCamera[i]; c.camera_to_world[i] = cam.M_CameratoWorld[i]; }
        c.lens_radius = cam.lensRadius; c.focal_distance = cam.focalDistance; c.camera_kind = cam.kind;
        c.sampler_kind = s.kind; c.xs = s.xs; c.ys = s.ys; c.jitter = s.jitter; c.seed = s.seed;
        c.filter_kind = f.kind; c.filter_rx = f.rx; c.filter_ry = f.ry; c.filter_sigma = f.sigma;
        c.mode = mode; c.max_depth = max_depth; c.rr_depth = rr_depth; c.ray_eps = ray_eps; c.shadow_eps = shadow_eps;
        for (int i = 0; i < 3; ++i) c.albedo[i] = albedo[i];
        c.spp_begin = spp_begin; c.spp_end = spp_end; c.rank = rank; c.world = world; c.partition = partition; c.tile_w = tile_w; c.tile_h = tile_h;
        return c;
    }
    // evaluate_pixel for every pixel and sample index in [spp_begin, spp_end) (RayTracerTestApp.h:287-409)
    crt_render_stats Render(Scene& scene, Film& film, const CameraBase& cam, const SamplerDesc& s, const FilterDesc& f, int spp_begin, int spp_end) const {
        crt_render_config c = Config(film, cam, s, f, spp_begin, spp_end);
        crt_render_stats st{};
        check(crt_render(scene.handle(), film.handle(), &c, &st));
        return st;
    }
};

}  // namespace crt
