// ORACLE SUPPORT (test infrastructure, never shipped, never on the product path).
//
// extern "C" probes over the REFERENCE'S OWN code, compiled unmodified from where it lies under
// /root/reference (recipe: oracle/Makefile target `ref`, output oracle/_ref/libcrt_ref.so).  This file
// contains no restatement of the hot path: every ref_* entry point constructs the reference's classes
// (Triangle, TriModel, Octtree_Model, Sphere/Cylinder/Disk/TriangleSimple, Perspective/Orthographic/
// PinholeCamera, pbrt::RNG / samplers / spectra / PixelSensor / filters) and returns what they compute, with
// the same argument layout as the matching orc_* probe in oracle_capi.cpp, so tests/test_cpu_ref_pin.py can
// call both and demand equality.  The only glue that is NOT reference code is marked GLUE below: the two
// lambdas of Applications/RayTracerTestApp.h (Li :218-284, evaluate_pixel :287-345) live inside an application
// function full of GL calls and cannot be included, so ref_eval_samples / ref_render_tier_a call the same
// reference functions in the same order as those lambdas.
//
// Third-party boundary: glm is not in the repository; oracle/refshim/glm/glm.hpp stands in for it (see its
// header for what that does and does not pin).  MSVC-isms are handled in oracle/refshim/ref_prelude.h.
#include "RayTracer/Sampling.h"
#include "RayTracer/Octtree_Model.h"
#include "RayTracer/Cameras.h"
#include "RayTracer/Film.h"
#include "ThirdParty/pbrv4/rng.h"
#include "ThirdParty/pbrv4/hash.h"
#include "ThirdParty/pbrv4/samplers.h"

#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>

// AssetManager.cpp:6 defines this static; that file is the assimp loader and is not compiled.
std::unordered_map<std::string, MeshCache::Model> MeshCache::modelCache;

namespace {

glm::mat4 mat_from(const float* p) { glm::mat4 m; std::memcpy(&m[0].x, p, 64); return m; }
void mat_to(const glm::mat4& m, float* p) { std::memcpy(p, &m[0].x, 64); }
void mat3_to(const glm::mat3& m, float* p) { for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) p[3 * c + r] = m[c][r]; }
Ray ray_from(const float* r) { return Ray(glm::vec3(r[0], r[1], r[2]), glm::vec3(r[3], r[4], r[5])); }

// Film::pixel_sensor selection: empty = sensor_xyz (RayTracerTestApp.h:149); otherwise the measured-sensor constructor
// (pixelsensor.h:37-68) on named response curves, as the app's sensor_canon (:152-153)
std::string g_sensor_names[4];
float g_sensor_ratio = 0;
std::once_flag g_init_once;
void init_tables() {
    // RayTracerTestApp.h:137-139.  RGBToSpectrumTable::Init prints "couldnt open rgb2spec file" (color.cpp:160-163):
    // the table file is not in the repository; grey RGB never reads the table (color.cpp:35-37).
    std::call_once(g_init_once, [] {
        pbrt::Spectra::Init();
        pbrt::RGBToSpectrumTable::Init();
        pbrt::RGBColorSpace::Init();
    });
}

pbrt::PixelSensor make_sensor() {
    if (g_sensor_names[0].empty())
        return pbrt::PixelSensor(pbrt::RGBColorSpace::sRGB, pbrt::GetNamedSpectrum("stdillum-D65"), 1.0f / pbrt::CIE_Y_integral);
    return pbrt::PixelSensor(pbrt::GetNamedSpectrum(g_sensor_names[0]), pbrt::GetNamedSpectrum(g_sensor_names[1]), pbrt::GetNamedSpectrum(g_sensor_names[2]),
                             pbrt::RGBColorSpace::sRGB, pbrt::GetNamedSpectrum(g_sensor_names[3]), g_sensor_ratio);
}

struct RScene {
    std::string model_name;
    std::unique_ptr<TriModel> model;
    std::unique_ptr<Octtree_Model> oct;
    std::vector<std::unique_ptr<Shape>> shapes;
    glm::mat4 rigid;
    Triangle::vertex_available avail;
};
int g_scene_counter = 0;

std::unique_ptr<pbrt::Sampler> make_sampler(int kind, int xs, int ys, int jitter, int seed) {
    if (kind == 0) return std::make_unique<pbrt::IndependentSampler>(xs * ys, seed);
    return std::make_unique<pbrt::StratifiedSampler>(xs, ys, jitter != 0, seed);
}

// protected matrices of the reference cameras, exposed by derivation only
struct PerspProbe : PerspectiveCamera { using PerspectiveCamera::PerspectiveCamera; using CameraBase::M_RastertoCamera; using CameraBase::M_CameratoWorld; };
struct OrthoProbe : OrthographicCamera { using OrthographicCamera::OrthographicCamera; using CameraBase::M_RastertoCamera; using CameraBase::M_CameratoWorld; };
struct PinholeProbe : PinholeCamera { using PinholeCamera::PinholeCamera; using CameraBase::M_RastertoScreen; using CameraBase::M_CameratoWorld; };

std::unique_ptr<CameraBase> make_camera(int kind, float near_, float far_, float sw, float sh, float fov, const float* pos, const float* look,
                                        const float* right, const float* up, float resx, float resy, float lens_radius, float focal_distance) {
    glm::vec3 p(pos[0], pos[1], pos[2]), l(look[0], look[1], look[2]), r(right[0], right[1], right[2]), u(up[0], up[1], up[2]);
    if (kind == 0) return std::make_unique<PerspProbe>(near_, far_, sw, sh, fov, p, l, r, u, glm::vec2(resx, resy), lens_radius, focal_distance);
    if (kind == 1) return std::make_unique<OrthoProbe>(near_, far_, sw, sh, p, l, r, u, glm::vec2(resx, resy));
    return std::make_unique<PinholeProbe>(0.25f, glm::vec3(sw, sh, far_), p, l, r, u, glm::vec2(resx, resy));
}

}  // namespace

extern "C" {

const char* ref_describe() {
    return "reference sources compiled unmodified from /root/reference (RayTracer/{Shapes,Octtree_Model,Cameras,Sampling,Film}.h, "
           "ThirdParty/pbrv4/{spectrum,color,colorspace,pixelsensor,filters}.cpp, ThirdParty/AABB_triangle_Moller.h) against oracle/refshim";
}

// ---------------------------------------------------------------- integer / sampling probes ------
uint64_t ref_murmur64a(const unsigned char* key, uint64_t len, uint64_t seed) { return pbrt::MurmurHash64A(key, (size_t)len, seed); }
uint64_t ref_mixbits(uint64_t v) { return pbrt::MixBits(v); }
uint64_t ref_helper_mixbits(uint64_t v) { return Helper::MixBits(v); }
uint64_t ref_hash_pixel_seed(int x, int y, int seed) { return pbrt::Hash(glm::ivec2(x, y), seed); }
uint64_t ref_hash_pixel_dim_seed(int x, int y, int dim, int seed) { return pbrt::Hash(glm::ivec2(x, y), dim, seed); }
int ref_permutation_element(uint32_t i, uint32_t l, uint32_t p) { return Helper::PermutationElement(i, l, p); }
void ref_pcg32(int mode, uint64_t seq, uint64_t offset, int64_t adv, int n, uint32_t* out_u32, float* out_f) {
    pbrt::RNG r;
    if (mode == 1) r.SetSequence(seq);
    if (mode == 2) r.SetSequence(seq, offset);
    if (adv) r.Advance(adv);
    for (int i = 0; i < n; ++i) {
        if (out_u32) out_u32[i] = r.Uniform<uint32_t>();
        else out_f[i] = r.Uniform<float>();
    }
}
void ref_sampler_sequence(int kind, int xs, int ys, int jitter, int seed, int px, int py, int index, int dim, const char* pattern, float* out) {
    auto s = make_sampler(kind, xs, ys, jitter, seed);
    s->StartPixelSample(glm::ivec2(px, py), index, dim);
    for (const char* c = pattern; *c; ++c) {
        if (*c == '1') *out++ = s->Get1D();
        else { glm::vec2 v = (*c == 'p') ? s->GetPixel2D() : s->Get2D(); *out++ = v.x; *out++ = v.y; }
    }
}
void ref_sample_visible(float u, float* lambda8, float* pdf8) {
    pbrt::SampledWavelengths w = pbrt::SampledWavelengths::SampleVisible(u);
    pbrt::SampledSpectrum pdf = w.PDF();
    for (int i = 0; i < 8; ++i) { lambda8[i] = w[i]; pdf8[i] = pdf[i]; }
}
// kind 0 only: TriangleFilter::Sample draws from a global mt19937 (Sampling.h:228-235) and is not a function of u
int ref_filter_sample(int kind, float rx, float ry, float u0, float u1, float* out3) {
    if (kind != 0) return -1;
    pbrt::FilterSample fs = pbrt::BoxFilter(glm::vec2(rx, ry)).Sample(glm::vec2(u0, u1));
    out3[0] = fs.p.x; out3[1] = fs.p.y; out3[2] = fs.weight;
    return 0;
}
// pbrt::GaussianFilter(radius, sigma).Sample(u) (filters.h:96-163) for n samples: (p.x, p.y, weight)
void ref_gaussian_filter_samples(float rx, float ry, float sigma, const float* u2, int n, float* out3) {
    pbrt::GaussianFilter f(glm::vec2(rx, ry), sigma);
    for (int i = 0; i < n; ++i) { pbrt::FilterSample fs = f.Sample(glm::vec2(u2[2 * i], u2[2 * i + 1])); out3[3 * i] = fs.p.x; out3[3 * i + 1] = fs.p.y; out3[3 * i + 2] = fs.weight; }
}
void ref_cosine_hemisphere(const float* u2, int n, float* w3, float* pdf) {
    for (int i = 0; i < n; ++i) { glm::vec3 w = SampleCosineHemisphere(glm::vec2(u2[2 * i], u2[2 * i + 1])); w3[3 * i] = w.x; w3[3 * i + 1] = w.y; w3[3 * i + 2] = w.z; pdf[i] = CosineHemispherePDF(w.z); }
}
void ref_terminate_secondary(float u, float* pdf8) {
    pbrt::SampledWavelengths w = pbrt::SampledWavelengths::SampleVisible(u);
    w.TerminateSecondary();
    pbrt::SampledSpectrum pdf = w.PDF();
    for (int i = 0; i < 8; ++i) pdf8[i] = pdf[i];
}
// SampleLinear(u, a, b) (RayTracer/Sampling.h:205-211): the two deterministic halves of SampleTent -- (0,1) and (1,0) -- given its coin
float ref_sample_linear(float u, float a, float b) { return SampleLinear(u, a, b); }
void ref_concentric_disk(float u0, float u1, float* out2) { glm::vec2 d = SampleUniformDiskConcentric(glm::vec2(u0, u1)); out2[0] = d.x; out2[1] = d.y; }
float ref_gamma(int n) { return pbrt::gamma(n); }
float ref_difference_of_products(float a, float b, float c, float d) { return pbrt::DifferenceOfProducts(a, b, c, d); }

// ---------------------------------------------------------------- spectra / colour ----------------
// which: 0 X, 1 Y, 2 Z, 3 the sRGB colour space's illuminant (stdillum-D65), queried at the 471 integer wavelengths
void ref_dense_table(int which, float* out471) {
    init_tables();
    for (int i = 0; i < 471; ++i) {
        float l = 360.0f + i;
        out471[i] = which == 0 ? pbrt::Spectra::X().Query(l) : which == 1 ? pbrt::Spectra::Y().Query(l) : which == 2 ? pbrt::Spectra::Z().Query(l)
                  : pbrt::RGBColorSpace::sRGB->illuminant.Query(l);
    }
}
// GetNamedSpectrum(name)->Query at n wavelengths; returns -1 for an unknown name
int ref_named_spectrum_query(const char* name, const float* lambdas, int n, float* out) {
    init_tables();
    pbrt::Spectrum* s = pbrt::GetNamedSpectrum(name);
    if (!s) return -1;
    for (int i = 0; i < n; ++i) out[i] = s->Query(lambdas[i]);
    return 0;
}
// PiecewiseLinearSpectrum::FromInterleaved(samples, normalize) queried at n wavelengths
void ref_interleaved_spectrum_query(const float* interleaved, int count, int normalize, const float* lambdas, int n, float* out) {
    init_tables();
    std::unique_ptr<pbrt::PiecewiseLinearSpectrum> s(pbrt::PiecewiseLinearSpectrum::FromInterleaved(std::span<const float>(interleaved, (size_t)count), normalize != 0));
    for (int i = 0; i < n; ++i) out[i] = s->Query(lambdas[i]);
}
void ref_color_constants(float* sensor9, float* rgbfromxyz9, float* xyzfromrgb9, float* white2) {
    init_tables();
    pbrt::PixelSensor sensor(pbrt::RGBColorSpace::sRGB, pbrt::GetNamedSpectrum("stdillum-D65"), 1.0f / pbrt::CIE_Y_integral);  // RayTracerTestApp.h:149
    mat3_to(sensor.XYZFromSensorRGB, sensor9);
    mat3_to(pbrt::RGBColorSpace::sRGB->RGBFromXYZ, rgbfromxyz9);
    mat3_to(pbrt::RGBColorSpace::sRGB->XYZFromRGB, xyzfromrgb9);
    white2[0] = pbrt::RGBColorSpace::sRGB->w.x; white2[1] = pbrt::RGBColorSpace::sRGB->w.y;
}
// Hands the reference a table the way its own Init does after reading the (absent) file: new RGBToSpectrumTable(scale, data)
// (color.cpp:164), then RGBColorSpace::Init() again so that sRGB captures the new table pointer (colorspace.cpp:89-90).
void ref_set_rgb_table(const float* scale64, const float* data) {
    init_tables();
    float* sc = new float[64];
    float* dt = new float[(size_t)3 * 64 * 64 * 64 * 3];
    std::memcpy(sc, scale64, 64 * sizeof(float));
    std::memcpy(dt, data, (size_t)3 * 64 * 64 * 64 * 3 * sizeof(float));
    pbrt::RGBToSpectrumTable::sRGB = new pbrt::RGBToSpectrumTable(sc, dt);
    pbrt::RGBColorSpace::Init();
}
// RGBColorSpace::ToRGBCoeffs(rgb) (colorspace.cpp:38-43), read back through the polynomial: p(0) = c2, and c0, c1 from p(1), p(-1)
// would lose bits -- so the spectrum itself is returned instead: RGBAlbedoSpectrum(sRGB, rgb).Query at n wavelengths
void ref_rgb_albedo_query(const float* rgb3, const float* lambdas, int n, float* out) {
    init_tables();
    pbrt::RGBAlbedoSpectrum s(*pbrt::RGBColorSpace::sRGB, pbrt::RGB(rgb3[0], rgb3[1], rgb3[2]));
    for (int i = 0; i < n; ++i) out[i] = s.Query(lambdas[i]);
}
// kind 0 RGBAlbedoSpectrum, 1 RGBIlluminantSpectrum, 2 RGBUnboundedSpectrum of rgb: Sample(SampleVisible(u))
void ref_rgb_spectrum_sample(int kind, const float* rgb3, float u, float* lambda8, float* out8) {
    init_tables();
    pbrt::SampledWavelengths w = pbrt::SampledWavelengths::SampleVisible(u);
    pbrt::RGB rgb(rgb3[0], rgb3[1], rgb3[2]);
    pbrt::SampledSpectrum s = kind == 0 ? pbrt::RGBAlbedoSpectrum(*pbrt::RGBColorSpace::sRGB, rgb).Sample(w)
                            : kind == 1 ? pbrt::RGBIlluminantSpectrum(*pbrt::RGBColorSpace::sRGB, rgb).Sample(w)
                                        : pbrt::RGBUnboundedSpectrum(*pbrt::RGBColorSpace::sRGB, rgb).Sample(w);
    for (int i = 0; i < 8; ++i) { lambda8[i] = w[i]; out8[i] = s[i]; }
}
// Select the film's sensor: names of the r, g, b response spectra and of the sensor illuminant in the reference's registry
// (spectrum.cpp:2733-2854), or r_name == NULL for the XYZ sensor.  Exports what an independent implementation needs as INPUT
// (the curves and the illuminant at the 471 integer wavelengths) and the resulting XYZFromSensorRGB.  Returns -1 for unknown names.
int ref_set_sensor(const char* r_name, const char* g_name, const char* b_name, const char* illum_name, float imaging_ratio,
                   float* curves3x471_out, float* illum471_out, float* matrix9_out) {
    init_tables();
    if (!r_name) { for (auto& n : g_sensor_names) n.clear(); return 0; }
    const char* names[4] = {r_name, g_name, b_name, illum_name};
    pbrt::Spectrum* sp[4];
    for (int i = 0; i < 4; ++i) { sp[i] = pbrt::GetNamedSpectrum(names[i]); if (!sp[i]) return -1; }
    for (int i = 0; i < 4; ++i) g_sensor_names[i] = names[i];
    g_sensor_ratio = imaging_ratio;
    for (int l = 0; l < 471; ++l) {
        if (curves3x471_out) for (int c = 0; c < 3; ++c) curves3x471_out[471 * c + l] = sp[c]->Query(360.0f + l);
        if (illum471_out) illum471_out[l] = sp[3]->Query(360.0f + l);
    }
    if (matrix9_out) mat3_to(make_sensor().XYZFromSensorRGB, matrix9_out);
    return 0;
}
// the selected sensor's ToSensorRGB(L, SampleVisible(u))
void ref_sensor_rgb(float u, const float* L8, float* rgb3) {
    init_tables();
    pbrt::PixelSensor sensor = make_sensor();
    pbrt::SampledWavelengths w = pbrt::SampledWavelengths::SampleVisible(u);
    pbrt::SampledSpectrum L(0);
    for (int i = 0; i < 8; ++i) L[i] = L8[i];
    pbrt::RGB c = sensor.ToSensorRGB(L, w);
    rgb3[0] = c.r; rgb3[1] = c.g; rgb3[2] = c.b;
}
float ref_sigmoid_eval(float c0, float c1, float c2, float lambda) { return pbrt::RGBSigmoidPolynomial(c0, c1, c2)(lambda); }
// grey RGBAlbedoSpectrum (kind 0) / RGBIlluminantSpectrum (kind 1) sampled at 8 wavelengths, as Li builds them (:246,:255)
void ref_grey_rgb_spectrum_sample(int kind, float g, float u, float* lambda8, float* out8) {
    init_tables();
    pbrt::SampledWavelengths w = pbrt::SampledWavelengths::SampleVisible(u);
    pbrt::SampledSpectrum s = kind == 0 ? pbrt::RGBAlbedoSpectrum(*pbrt::RGBColorSpace::sRGB, pbrt::RGB(g, g, g)).Sample(w)
                                        : pbrt::RGBIlluminantSpectrum(*pbrt::RGBColorSpace::sRGB, pbrt::RGB(g, g, g)).Sample(w);
    for (int i = 0; i < 8; ++i) { lambda8[i] = w[i]; out8[i] = s[i]; }
}
// PixelSensor::ToSensorRGB(L, lambdas) for lambdas = SampleVisible(u)
void ref_to_sensor_rgb(float u, const float* L8, float* rgb3) {
    init_tables();
    pbrt::PixelSensor sensor(pbrt::RGBColorSpace::sRGB, pbrt::GetNamedSpectrum("stdillum-D65"), 1.0f / pbrt::CIE_Y_integral);
    pbrt::SampledWavelengths w = pbrt::SampledWavelengths::SampleVisible(u);
    pbrt::SampledSpectrum L(0);
    for (int i = 0; i < 8; ++i) L[i] = L8[i];
    pbrt::RGB c = sensor.ToSensorRGB(L, w);
    rgb3[0] = c.r; rgb3[1] = c.g; rgb3[2] = c.b;
}

// ---------------------------------------------------------------- cameras / transforms ------------
void ref_camera_matrices(int kind, float near_, float far_, float sw, float sh, float fov, const float* pos, const float* look,
                         const float* right, const float* up, float resx, float resy, float* r2c16, float* c2w16) {
    auto cam = make_camera(kind, near_, far_, sw, sh, fov, pos, look, right, up, resx, resy, 0, 0);
    if (kind == 0) { auto* c = static_cast<PerspProbe*>(cam.get()); mat_to(c->M_RastertoCamera, r2c16); mat_to(c->M_CameratoWorld, c2w16); }
    else if (kind == 1) { auto* c = static_cast<OrthoProbe*>(cam.get()); mat_to(c->M_RastertoCamera, r2c16); mat_to(c->M_CameratoWorld, c2w16); }
    else { auto* c = static_cast<PinholeProbe*>(cam.get()); mat_to(c->M_RastertoScreen, r2c16); mat_to(c->M_CameratoWorld, c2w16); }
}
// generateRay for n film positions; the lens sample (if any) comes from a StratifiedSampler(xs, ys, jitter, seed)
// started at (pixel = floor(film pos), index, dim) — i.e. the state evaluate_pixel would hand over at that dimension
void ref_camera_rays(int kind, float near_, float far_, float sw, float sh, float fov, const float* pos, const float* look, const float* right,
                     const float* up, float resx, float resy, float lens_radius, float focal_distance, const float* film_xy, int n,
                     int xs, int ys, int jitter, int seed, int index, int dim, float* rays6) {
    auto cam = make_camera(kind, near_, far_, sw, sh, fov, pos, look, right, up, resx, resy, lens_radius, focal_distance);
    pbrt::StratifiedSampler sampler(xs, ys, jitter != 0, seed);
    for (int i = 0; i < n; ++i) {
        sampler.StartPixelSample(glm::ivec2((int)film_xy[2 * i], (int)film_xy[2 * i + 1]), index, dim);
        Ray r = cam->generateRay(glm::vec2(film_xy[2 * i], film_xy[2 * i + 1]), &sampler);
        for (int k = 0; k < 3; ++k) { rays6[6 * i + k] = r.o[k]; rays6[6 * i + 3 + k] = r.d[k]; }
    }
}
void ref_shape_matrices(const float* rigid16, float* o2r16, float* r2o16) {
    Sphere s("m", mat_from(rigid16), 1, -1, 1, 360);
    mat_to(s.GetObjectToRenderMatrix(), o2r16);
    mat_to(s.GetRenderToObjectMatrix(), r2o16);
}

// ---------------------------------------------------------------- scene --------------------------
void* ref_scene_create() {
    init_tables();
    auto* s = new RScene;
    s->model_name = "ref_scene_" + std::to_string(g_scene_counter++);
    return s;
}
void ref_scene_destroy(void* h) {
    auto* s = (RScene*)h;
    s->oct.reset(); s->model.reset();
    MeshCache::modelCache.erase(s->model_name);
    delete s;
}
int ref_scene_set_model(void* h, int n_meshes, const float* positions, const float* normals, const uint32_t* nverts,
                        const uint32_t* indices, const uint32_t* ntris, const float* rigid16, int precomputed_world,
                        int cull_backface, const float* look_dir, const float* texcoords, const float* tangents, const float* bitangents) {
    auto* s = (RScene*)h;
    MeshCache::Model model;
    model.mesh_name = s->model_name;
    size_t vo = 0, io = 0;
    for (int m = 0; m < n_meshes; ++m) {
        MeshCache::Mesh mesh;
        for (uint32_t v = 0; v < nverts[m]; ++v) {
            mesh.positions.push_back(glm::vec3(positions[3 * (vo + v)], positions[3 * (vo + v) + 1], positions[3 * (vo + v) + 2]));
            if (normals) mesh.normals.push_back(glm::vec3(normals[3 * (vo + v)], normals[3 * (vo + v) + 1], normals[3 * (vo + v) + 2]));
        }
        // CalculateLocalSurface reads texcoords/tangents/bitangents/normals of the three vertices unconditionally and only
        // then looks at vertex_available (Shapes.h:1040-1071): the arrays must exist even when flagged unavailable (the
        // assimp loader always fills them).  Zero-filled here; their values are never used with the flags below.
        mesh.texcoords.assign(nverts[m], glm::vec2(0, 0));
        mesh.tangents.assign(nverts[m], glm::vec3(0, 0, 0));
        mesh.bitangents.assign(nverts[m], glm::vec3(0, 0, 0));
        if (!normals) mesh.normals.assign(nverts[m], glm::vec3(0, 0, 0));
        for (uint32_t v = 0; v < nverts[m]; ++v) {
            if (texcoords) mesh.texcoords[v] = glm::vec2(texcoords[2 * (vo + v)], texcoords[2 * (vo + v) + 1]);
            if (tangents) mesh.tangents[v] = glm::vec3(tangents[3 * (vo + v)], tangents[3 * (vo + v) + 1], tangents[3 * (vo + v) + 2]);
            if (bitangents) mesh.bitangents[v] = glm::vec3(bitangents[3 * (vo + v)], bitangents[3 * (vo + v) + 1], bitangents[3 * (vo + v) + 2]);
        }
        mesh.indices.assign(indices + io, indices + io + 3 * (size_t)ntris[m]);
        vo += nverts[m]; io += 3 * (size_t)ntris[m];
        model.meshes.push_back(std::move(mesh));
    }
    MeshCache::modelCache[s->model_name] = std::move(model);
    Triangle::vertex_available avail;
    avail.texcoords = texcoords != nullptr; avail.tangents = tangents != nullptr; avail.bitangents = bitangents != nullptr;
    avail.normals = normals != nullptr;
    avail.precomputed_worldtransform = precomputed_world != 0;
    s->avail = avail;
    s->rigid = mat_from(rigid16);
    s->model = std::make_unique<TriModel>("model", s->rigid, s->model_name, cull_backface != 0, precomputed_world != 0, avail);
    if (cull_backface && normals) s->model->ComputeBackFace(glm::vec3(look_dir[0], look_dir[1], look_dir[2]), true);
    return 0;
}
int ref_scene_build_octree(void* h) {
    auto* s = (RScene*)h;
    s->oct = std::make_unique<Octtree_Model>(*s->model);
    s->oct->CreateOcttree();
    return s->oct->getTreeSize();
}
// same layout as orc_octree_dump, read through the public GetNode(i) (Octtree_Model.h:178)
long long ref_octree_dump(void* h, float* bounds6, int32_t* leaf, int32_t* child8, long long* list_off, int32_t* list_pairs, long long cap_pairs) {
    auto* s = (RScene*)h;
    long long off = 0;
    int nn = s->oct->getTreeSize();
    for (int i = 0; i < nn; ++i) {
        Octtree_Model::node n = s->oct->GetNode(i);
        if (bounds6) { bounds6[6 * i] = n.bounds.pmin.x; bounds6[6 * i + 1] = n.bounds.pmin.y; bounds6[6 * i + 2] = n.bounds.pmin.z;
                       bounds6[6 * i + 3] = n.bounds.pmax.x; bounds6[6 * i + 4] = n.bounds.pmax.y; bounds6[6 * i + 5] = n.bounds.pmax.z; }
        if (leaf) leaf[i] = n.leaf ? 1 : 0;
        if (child8) for (int k = 0; k < 8; ++k) child8[8 * i + k] = n.leaf ? -1 : n.child_id[k];
        if (list_off) list_off[i] = off;
        for (const auto& ii : n.triangle_info) {
            if (list_pairs && off < cap_pairs) { list_pairs[2 * off] = ii.mesh_id; list_pairs[2 * off + 1] = ii.tri_id; }
            ++off;
        }
    }
    if (list_off) list_off[nn] = off;
    return off;
}
// raw pointers to the reference objects, for the private-member readers in ref_harness_private.cpp
const void* ref_scene_tri_model(void* h) { return ((RScene*)h)->model.get(); }
const void* ref_scene_octree(void* h) { return ((RScene*)h)->oct.get(); }
void ref_model_bounds(void* h, float* out6) {
    Bounds3 b = ((RScene*)h)->model->Bounds();
    out6[0] = b.pmin.x; out6[1] = b.pmin.y; out6[2] = b.pmin.z; out6[3] = b.pmax.x; out6[4] = b.pmax.y; out6[5] = b.pmax.z;
}
int ref_scene_add_shape(void* h, int kind, const float* rigid16, const float* params) {
    auto* s = (RScene*)h;
    glm::mat4 M = mat_from(rigid16);
    std::unique_ptr<Shape> sh;
    if (kind == 0) sh = std::make_unique<Sphere>("s", M, params[0], params[1], params[2], params[3]);
    else if (kind == 1) sh = std::make_unique<Cylinder>("c", M, params[0], params[1], params[2], params[3]);
    else if (kind == 2) sh = std::make_unique<Disk>("d", M, params[0], params[1], params[2], params[3]);
    else if (kind == 3) sh = std::make_unique<TriangleSimple>("t", M, glm::vec3(params[0], params[1], params[2]), glm::vec3(params[3], params[4], params[5]), glm::vec3(params[6], params[7], params[8]));
    else return -1;
    s->shapes.push_back(std::move(sh));
    return (int)s->shapes.size() - 1;
}

// ---------------------------------------------------------------- intersection probes -------------
// Bounds3::IntersectP(ray, tMax) (Shapes.h:100-124) for n (box, ray, tMax) triples
void ref_slab_test(const float* boxes6, const float* rays, const float* tmax, int n, int32_t* hit) {
    for (int i = 0; i < n; ++i) {
        Bounds3 b(glm::vec3(boxes6[6 * i], boxes6[6 * i + 1], boxes6[6 * i + 2]), glm::vec3(boxes6[6 * i + 3], boxes6[6 * i + 4], boxes6[6 * i + 5]));
        hit[i] = b.IntersectP(ray_from(rays + 6 * i), tmax[i]) ? 1 : 0;
    }
}
// Triangle::BasicIntersect(ray, tMax) (Shapes.h:1101-1260) of triangle (mesh_id[i], tri_id[i]) for ray i
void ref_triangle_intersect(void* h, const int32_t* mesh_id, const int32_t* tri_id, const float* rays, const float* tmax, int n,
                            int32_t* found, float* t, float* bary3) {
    auto* s = (RScene*)h;
    for (int i = 0; i < n; ++i) {
        Triangle tri("tri", s->rigid, s->model_name, mesh_id[i], tri_id[i], s->avail);
        auto r = tri.BasicIntersect(ray_from(rays + 6 * i), tmax[i]);
        found[i] = r.has_value();
        if (r) { t[i] = r->t; bary3[3 * i] = r->b0; bary3[3 * i + 1] = r->b1; bary3[3 * i + 2] = r->b2; }
    }
}
// TriModel::BasicIntersect (brute force, Shapes.h:1414-1471): the reference's own un-accelerated closest hit, WITH ids
void ref_brute_force(void* h, const float* rays, int n, int32_t* mesh_id, int32_t* tri_id, float* t, float* bary3) {
    auto* s = (RScene*)h;
    for (int i = 0; i < n; ++i) {
        auto r = s->model->BasicIntersect(ray_from(rays + 6 * i));
        mesh_id[i] = r ? r->mesh_id : -1; tri_id[i] = r ? r->tri_id : -1;
        t[i] = r ? r->tri_isect.t : 0;
        if (r) { bary3[3 * i] = r->tri_isect.b0; bary3[3 * i + 1] = r->tri_isect.b1; bary3[3 * i + 2] = r->tri_isect.b2; }
    }
}
// Octtree_Model::Traverse (Octtree_Model.h:66-127): found, tHit, n, hitp, u, v.  The hit id is a local of Traverse and
// is not observable; ref_surface_of reproduces the returned record from a claimed id instead.
void ref_traverse_surface(void* h, const float* rays, int n, int nthreads, int32_t* found, float* thit, float* nrm3, float* hitp3, float* uv2) {
    auto* s = (RScene*)h;
    auto work = [&](int b, int e) {
        for (int i = b; i < e; ++i) {
            Ray ray = ray_from(rays + 6 * i);
            auto r = s->oct->Traverse(ray);
            found[i] = r.has_value();
            if (r) { thit[i] = r->tHit; for (int k = 0; k < 3; ++k) { nrm3[3 * i + k] = r->n[k]; hitp3[3 * i + k] = r->hitp[k]; } uv2[2 * i] = r->u; uv2[2 * i + 1] = r->v; }
        }
    };
    nthreads = std::max(1, std::min(nthreads, n));
    std::vector<std::thread> pool;
    int per = n / nthreads, b = 0;
    for (int t = 0; t < nthreads; ++t) { int e = (t == nthreads - 1) ? n : b + per; pool.emplace_back(work, b, e); b = e; }
    for (auto& th : pool) th.join();
}
// For a CLAIMED hit id: Triangle(mesh, tri).BasicIntersect(ray, FLT_MAX) -> CalculateLocalSurface, same record layout
void ref_surface_of(void* h, const int32_t* mesh_id, const int32_t* tri_id, const float* rays, int n, int32_t* found, float* thit, float* nrm3, float* hitp3, float* uv2) {
    auto* s = (RScene*)h;
    for (int i = 0; i < n; ++i) {
        found[i] = 0;
        if (mesh_id[i] < 0) continue;
        Triangle tri("tri", s->rigid, s->model_name, mesh_id[i], tri_id[i], s->avail);
        auto is = tri.BasicIntersect(ray_from(rays + 6 * i));
        if (!is) continue;
        auto r = tri.CalculateLocalSurface(*is);
        found[i] = r.has_value();
        if (r) { thit[i] = r->tHit; for (int k = 0; k < 3; ++k) { nrm3[3 * i + k] = r->n[k]; hitp3[3 * i + k] = r->hitp[k]; } uv2[2 * i] = r->u; uv2[2 * i + 1] = r->v; }
    }
}
static void put_info17(const LocalSurfaceInfo& r, float* w) {
    for (int k = 0; k < 3; ++k) { w[k] = r.hitp[k]; w[5 + k] = r.du[k]; w[8 + k] = r.dv[k]; w[11 + k] = r.n[k]; w[14 + k] = r.wo[k]; }
    w[3] = r.u; w[4] = r.v;
}
// Octtree_Model::Traverse, whole LocalSurfaceInfo record (Shapes.h:144-170) minus the never-assigned tHit: 17 floats per ray
void ref_traverse_local_surface(void* h, const float* rays, int n, int32_t* found, float* info17) {
    auto* s = (RScene*)h;
    for (int i = 0; i < n; ++i) {
        Ray ray = ray_from(rays + 6 * i);
        auto r = s->oct->Traverse(ray);
        found[i] = r.has_value();
        if (r) put_info17(*r, info17 + 17 * (size_t)i);
    }
}
// Triangle(mesh, tri).CalculateLocalSurface for given barycentrics and (normalised) ray direction, bypassing BasicIntersect
// (TriangleIntersect is a private type of Triangle: the braced list names no type)
void ref_local_surface_of(void* h, const int32_t* mesh_id, const int32_t* tri_id, const float* bary3, const float* rayd3, int n, float* info17) {
    auto* s = (RScene*)h;
    for (int i = 0; i < n; ++i) {
        Triangle tri("tri", s->rigid, s->model_name, mesh_id[i], tri_id[i], s->avail);
        auto r = tri.CalculateLocalSurface({bary3[3 * i], bary3[3 * i + 1], bary3[3 * i + 2], 0.0f, glm::vec3(rayd3[3 * i], rayd3[3 * i + 1], rayd3[3 * i + 2])});
        if (r) put_info17(*r, info17 + 17 * (size_t)i);
    }
}
// Shape::Intersect, the rest of the record: du (3), dv (3), wo (3) per ray (zero where nothing was hit)
void ref_shape_frame(void* h, int shape, const float* rays, int n, float tmax, float* du_dv_wo9) {
    auto* s = (RScene*)h;
    for (int i = 0; i < n; ++i) {
        Ray ray = ray_from(rays + 6 * i);
        auto r = s->shapes[shape]->Intersect(ray, tmax);
        if (r) for (int k = 0; k < 3; ++k) { du_dv_wo9[9 * i + k] = r->du[k]; du_dv_wo9[9 * i + 3 + k] = r->dv[k]; du_dv_wo9[9 * i + 6 + k] = r->wo[k]; }
    }
}
float ref_shape_area(void* h, int shape) { return ((RScene*)h)->shapes[shape]->Area(); }
void ref_triangle_area(void* h, const int32_t* mesh_id, const int32_t* tri_id, int n, float* out) {
    auto* s = (RScene*)h;
    for (int i = 0; i < n; ++i) out[i] = Triangle("tri", s->rigid, s->model_name, mesh_id[i], tri_id[i], s->avail).Area();
}
void ref_shape_intersect(void* h, int shape, const float* rays, int n, float tmax, int32_t* found, float* t, float* hitp3, float* nrm3, float* uv2) {
    auto* s = (RScene*)h;
    for (int i = 0; i < n; ++i) {
        auto r = s->shapes[shape]->Intersect(ray_from(rays + 6 * i), tmax);
        found[i] = r.has_value();
        if (r) { t[i] = r->tHit; for (int k = 0; k < 3; ++k) { hitp3[3 * i + k] = r->hitp[k]; nrm3[3 * i + k] = r->n[k]; } uv2[2 * i] = r->u; uv2[2 * i + 1] = r->v; }
    }
}
// Moller::triBoxOverlap (AABB_triangle_Moller.h:229-474) on n (center, half, triangle) triples
void ref_tri_box_overlap(const float* center3, const float* half3, const float* tri9, int n, int32_t* out) {
    for (int i = 0; i < n; ++i) {
        glm::vec3 c(center3[3 * i], center3[3 * i + 1], center3[3 * i + 2]);
        glm::vec3 hs(half3[3 * i], half3[3 * i + 1], half3[3 * i + 2]);
        std::array<glm::vec3, 3> tv;
        for (int a = 0; a < 3; ++a) tv[a] = glm::vec3(tri9[9 * i + 3 * a], tri9[9 * i + 3 * a + 1], tri9[9 * i + 3 * a + 2]);
        out[i] = Moller::triBoxOverlap(c, hs, tv);
    }
}

// ---------------------------------------------------------------- Tier A render (GLUE) ------------
struct ref_render_params {
    int width, height;
    int camera_kind;
    float near_, far_, sensor_w, sensor_h, fov;
    float pos[3], look[3], right[3], up[3];
    float lens_radius, focal_distance;
    int sampler_kind, xs, ys, jitter, seed;
    float filter_rx, filter_ry;   // filter radius
    float albedo[3];              // `colors` of RayTracerTestApp.h:208 (grey only: the RGB table file is absent)
    int spp_begin, spp_end, nthreads;
    int pixel_stride;             // >= 1: only pixel ids that are multiples of it are rendered (bounded samples for timing)
    int filter_kind;              // 0 BoxFilter, 2 GaussianFilter(radius, filter_sigma)  (1 = TriangleFilter is non-deterministic: refused)
    float filter_sigma;
};

namespace {
struct RenderSetup {
    std::unique_ptr<CameraBase> cam;
    std::unique_ptr<pbrt::Filter> filter;
    pbrt::PixelSensor sensor;
    Film film;
    pbrt::Spectrum* illumF;
    Octtree_Model* oct;
    float colors[3];
    RenderSetup(RScene* s, const ref_render_params* p)
        : cam(make_camera(p->camera_kind, p->near_, p->far_, p->sensor_w, p->sensor_h, p->fov, p->pos, p->look, p->right, p->up,
                          (float)p->width, (float)p->height, p->lens_radius, p->focal_distance)),
          filter(p->filter_kind == 2 ? std::unique_ptr<pbrt::Filter>(new pbrt::GaussianFilter(glm::vec2(p->filter_rx, p->filter_ry), p->filter_sigma > 0 ? p->filter_sigma : 0.5f))
                                     : std::unique_ptr<pbrt::Filter>(new pbrt::BoxFilter(glm::vec2(p->filter_rx, p->filter_ry)))),
          sensor(make_sensor()),                                                                                       // :149 / :152
          illumF(pbrt::GetNamedSpectrum("stdillum-F1")),                                                               // :196
          oct(s->oct.get()) {
        film.film_dim = glm::ivec2(p->width, p->height);      // :157-161
        film.image_res = glm::ivec2(p->width, p->height);
        film.filter = filter.get();
        film.pixel_sensor = &sensor;
        film.pixels.resize((size_t)p->width * p->height);
        for (int i = 0; i < 3; ++i) colors[i] = p->albedo[i];
    }
    // GLUE: RayTracerTestApp.h:218-284 with the dead `if (false)` branches removed; every call is reference code
    pbrt::SampledSpectrum Li(Ray ray, pbrt::SampledWavelengths lambdas) {
        std::optional<LocalSurfaceInfo> surfaceoptional = oct->Traverse(ray);
        if (surfaceoptional.has_value()) {
            LocalSurfaceInfo surfaceinfo = surfaceoptional.value();
            glm::vec3 world_n = surfaceinfo.n;
            pbrt::SampledSpectrum radiance = pbrt::SampledSpectrum(0);
            pbrt::RGBIlluminantSpectrum light_spectrum = pbrt::RGBIlluminantSpectrum(*pbrt::RGBColorSpace::sRGB, pbrt::RGB(1, 1, 1));
            pbrt::SampledSpectrum light_spectral = light_spectrum.Sample(lambdas);
            pbrt::SampledSpectrum ambient_spectral = 0.3f * illumF->Sample(lambdas);
            pbrt::RGBAlbedoSpectrum mat_spectrum(*pbrt::RGBColorSpace::sRGB, pbrt::RGB(colors[0], colors[1], colors[2]));
            pbrt::SampledSpectrum mat_spectral = mat_spectrum.Sample(lambdas);
            float light_1_cos = glm::clamp(glm::dot(world_n, glm::vec3(0, 0, -1)), 0.0f, 1.0f);
            radiance += ambient_spectral;
            radiance += light_1_cos * (light_spectral * mat_spectral);
            return radiance;
        }
        return pbrt::SampledSpectrum(0);
    }
    struct Debug { Ray ray{glm::vec3(0), glm::vec3(0)}; float lambda[8], pdf[8], L[8], rgb[3], weight; };
    // GLUE: RayTracerTestApp.h:287-345 (geometry_test == false)
    void evaluate_pixel(int pixel_id, int index, pbrt::Sampler* sampl, Debug* dbg) {
        int x_pix = pixel_id % film.image_res.x;
        int y_pix = film.image_res.y - std::floor(pixel_id / (float)film.image_res.x);
        glm::ivec2 pixel(x_pix, y_pix);
        sampl->StartPixelSample(pixel, index, 0);
        pbrt::SampledWavelengths lambdas = pbrt::SampledWavelengths::SampleVisible(sampl->Get1D());
        glm::vec2 uniform_pixel_offset = sampl->GetPixel2D();
        pbrt::FilterSample fs = film.filter->Sample(uniform_pixel_offset);
        glm::vec2 pixel_sampled_pos = glm::fvec2(pixel) + glm::vec2(.5f, .5f) + fs.p;
        Ray ray = cam->generateRay(pixel_sampled_pos, sampl);
        pbrt::SampledSpectrum L = Li(ray, lambdas);
        pbrt::RGB cam_RGB = film.pixel_sensor->ToSensorRGB(L, lambdas);
        cam_RGB.r = glm::clamp(cam_RGB.r, 0.0f, 1.0f);
        cam_RGB.g = glm::clamp(cam_RGB.g, 0.0f, 1.0f);
        cam_RGB.b = glm::clamp(cam_RGB.b, 0.0f, 1.0f);
        if (dbg) {
            dbg->ray = ray;
            pbrt::SampledSpectrum pdf = lambdas.PDF();
            for (int k = 0; k < 8; ++k) { dbg->lambda[k] = lambdas[k]; dbg->pdf[k] = pdf[k]; dbg->L[k] = L[k]; }
            dbg->rgb[0] = cam_RGB.r; dbg->rgb[1] = cam_RGB.g; dbg->rgb[2] = cam_RGB.b; dbg->weight = fs.weight;
            return;
        }
        film.pixels[pixel_id].rgbsum += fs.weight * cam_RGB.getglm();
        film.pixels[pixel_id].weightsum += fs.weight;
    }
};
}  // namespace

void ref_eval_samples(void* h, const ref_render_params* p, const int32_t* pixel_ids, const int32_t* indices, int n, float* ray6, float* lambda8,
                      float* pdf8, float* L8, float* rgb3, float* weight) {
    RenderSetup c((RScene*)h, p);
    auto sampler = make_sampler(p->sampler_kind, p->xs, p->ys, p->jitter, p->seed);
    for (int i = 0; i < n; ++i) {
        RenderSetup::Debug d;
        c.evaluate_pixel(pixel_ids[i], indices[i], sampler.get(), &d);
        for (int k = 0; k < 3; ++k) { ray6[6 * i + k] = d.ray.o[k]; ray6[6 * i + 3 + k] = d.ray.d[k]; rgb3[3 * i + k] = d.rgb[k]; }
        for (int k = 0; k < 8; ++k) { lambda8[8 * i + k] = d.lambda[k]; pdf8[8 * i + k] = d.pdf[k]; L8[8 * i + k] = d.L[k]; }
        weight[i] = d.weight;
    }
}
// Threaded exactly as RayTracerTestApp.h:372-409: static contiguous pixel ranges, one sampler per thread, one
// sample index per pass.  film_io (rgbsum, weightsum per pixel) is accumulated into.
void ref_render_tier_a(void* h, const ref_render_params* p, float* film_io) {
    RenderSetup c((RScene*)h, p);
    size_t np = c.film.pixels.size();
    for (size_t i = 0; i < np; ++i) { c.film.pixels[i].rgbsum = glm::vec3(film_io[4 * i], film_io[4 * i + 1], film_io[4 * i + 2]); c.film.pixels[i].weightsum = film_io[4 * i + 3]; }
    int nthreads = std::max(1, std::min<int>(p->nthreads, (int)np));
    const size_t stride = (size_t)std::max(1, p->pixel_stride);
    std::vector<std::thread> pool;
    size_t per = np / nthreads, b = 0;
    for (int t = 0; t < nthreads; ++t) {
        size_t e = (t == nthreads - 1) ? np : b + per;
        pool.emplace_back([&, b, e] {
            auto sampler = make_sampler(p->sampler_kind, p->xs, p->ys, p->jitter, p->seed);
            for (int index = p->spp_begin; index < p->spp_end; ++index)
                for (size_t px = (b + stride - 1) / stride * stride; px < e; px += stride) c.evaluate_pixel((int)px, index, sampler.get(), nullptr);
        });
        b = e;
    }
    for (auto& th : pool) th.join();
    for (size_t i = 0; i < np; ++i) { film_io[4 * i] = c.film.pixels[i].rgbsum.x; film_io[4 * i + 1] = c.film.pixels[i].rgbsum.y; film_io[4 * i + 2] = c.film.pixels[i].rgbsum.z; film_io[4 * i + 3] = c.film.pixels[i].weightsum; }
}
// GLUE: film resolve, RayTracerTestApp.h:425-452 (geometry_test == false)
void ref_resolve(const float* film4, int npix, unsigned char* rgb8, float* rgbf) {
    init_tables();
    pbrt::PixelSensor sensor = make_sensor();
    for (int i = 0; i < npix; ++i) {
        glm::vec3 rgbsum(film4[4 * i], film4[4 * i + 1], film4[4 * i + 2]);
        pbrt::RGB sensor_rgb(rgbsum / film4[4 * i + 3]);
        pbrt::XYZ xyz_val = sensor.XYZFromSensorRGB * sensor_rgb.getglm();
        pbrt::RGB output_rgb = pbrt::RGBColorSpace::sRGB->ToRGB(xyz_val);
        output_rgb.r = glm::clamp(output_rgb.r, 0.0f, 1.0f);
        output_rgb.g = glm::clamp(output_rgb.g, 0.0f, 1.0f);
        output_rgb.b = glm::clamp(output_rgb.b, 0.0f, 1.0f);
        float scale = 255.0f;
        if (rgb8) { rgb8[3 * i] = scale * output_rgb.r; rgb8[3 * i + 1] = scale * output_rgb.g; rgb8[3 * i + 2] = scale * output_rgb.b; }
        if (rgbf) { rgbf[3 * i] = output_rgb.r; rgbf[3 * i + 1] = output_rgb.g; rgbf[3 * i + 2] = output_rgb.b; }
    }
}

}  // extern "C"
