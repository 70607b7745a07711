// ORACLE (test infrastructure, never shipped, never on the product path).
// extern "C" surface used by tests/ (ctypes), __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference arm.  Nothing under computational_ray_tracer_b200/ links or loads this library.
#include <cstdio>
#include <functional>

#include "oracle_render.h"

using namespace orc;

namespace orc {
const float* swatch_table(int i, int* n);
const float* named_table(const char* name, int* n);
}

namespace {
struct OScene {
    std::string model_name;
    std::unique_ptr<TriModel> model;
    std::unique_ptr<Octtree_Model> oct;
    std::vector<std::unique_ptr<Shape>> shapes;
    Scene scene;
};
int g_scene_counter = 0;
// Film::pixel_sensor.  Default: the app's sensor_xyz (RayTracerTestApp.h:149); orc_set_sensor swaps in a measured sensor
// (its sensor_canon, :152-153).
struct DenseFromArray : DenselySampledSpectrum { explicit DenseFromArray(const float* v) { values.assign(v, v + 471); } };
std::unique_ptr<PixelSensor> g_sensor_override;
std::unique_ptr<PixelSensor> make_sensor() {
    if (g_sensor_override) return std::make_unique<PixelSensor>(*g_sensor_override);
    return std::make_unique<PixelSensor>(RGBColorSpace::sRGB(), SpectraTables::get().illumD65.get(), 1.0f / CIE_Y_integral);
}
}  // namespace

extern "C" {

// ---------------------------------------------------------------- known-answer entry points ------
uint64_t orc_murmur64a(const unsigned char* key, uint64_t len, uint64_t seed) { return MurmurHash64A(key, (size_t)len, seed); }
uint64_t orc_mixbits(uint64_t v) { return MixBits(v); }
uint64_t orc_hash_pixel_seed(int x, int y, int seed) { return HashPixelSeed(ivec2(x, y), seed); }
uint64_t orc_hash_pixel_dim_seed(int x, int y, int dim, int seed) { return HashPixelDimSeed(ivec2(x, y), dim, seed); }
int orc_permutation_element(uint32_t i, uint32_t l, uint32_t p) { return PermutationElement(i, l, p); }
// mode 0: RNG() default state; 1: SetSequence(seq); 2: SetSequence(seq, offset).  Then Advance(adv), then n draws.
void orc_pcg32(int mode, uint64_t seq, uint64_t offset, int64_t adv, int n, uint32_t* out_u32, float* out_f) {
    RNG r;
    if (mode == 1) r.SetSequence(seq);
    if (mode == 2) r.SetSequence(seq, offset);
    if (adv) r.Advance(adv);
    for (int i = 0; i < n; ++i) {
        if (out_u32) out_u32[i] = r.UniformU32();
        else out_f[i] = r.UniformFloat();
    }
}
static std::unique_ptr<Sampler> make_sampler(int kind, int xs, int ys, int jitter, int seed) {
    if (kind == 0) return std::make_unique<IndependentSampler>(xs * ys, seed);
    return std::make_unique<StratifiedSampler>(xs, ys, jitter != 0, seed);
}
// pattern: string of '1' (Get1D), '2' (Get2D), 'p' (GetPixel2D); out gets 1 or 2 floats per char
void orc_sampler_sequence(int kind, int xs, int ys, int jitter, int seed, int px, int py, int index, int dim, const char* pattern, float* out) {
    auto s = make_sampler(kind, xs, ys, jitter, seed);
    s->StartPixelSample(ivec2(px, py), index, dim);
    for (const char* c = pattern; *c; ++c) {
        if (*c == '1') *out++ = s->Get1D();
        else { vec2 v = (*c == 'p') ? s->GetPixel2D() : s->Get2D(); *out++ = v.x; *out++ = v.y; }
    }
}
void orc_sample_visible(float u, float* lambda8, float* pdf8) {
    SampledWavelengths w = SampledWavelengths::SampleVisible(u);
    for (int i = 0; i < 8; ++i) { lambda8[i] = w.lambda[i]; pdf8[i] = w.pdf[i]; }
}
void orc_filter_sample(int kind, float rx, float ry, float u0, float u1, float* out3) {
    FilterSample fs = (kind == 0) ? BoxFilter(vec2(rx, ry)).Sample(vec2(u0, u1)) : TriangleFilter(vec2(rx, ry)).Sample(vec2(u0, u1));
    out3[0] = fs.p.x; out3[1] = fs.p.y; out3[2] = fs.weight;
}
// GaussianFilter(radius, sigma).Sample(u) for n samples: out = (p.x, p.y, weight) per sample (filters.h:96-163)
void orc_gaussian_filter_samples(float rx, float ry, float sigma, const float* u2, int n, float* out3) {
    GaussianFilter f(vec2(rx, ry), sigma);
    for (int i = 0; i < n; ++i) { FilterSample fs = f.Sample(vec2(u2[2 * i], u2[2 * i + 1])); out3[3 * i] = fs.p.x; out3[3 * i + 1] = fs.p.y; out3[3 * i + 2] = fs.weight; }
}
// Tier-B building blocks that DO exist in the reference: SampleCosineHemisphere / CosineHemispherePDF (Sampling.h:449-459)
void orc_cosine_hemisphere(const float* u2, int n, float* w3, float* pdf) {
    for (int i = 0; i < n; ++i) { vec3 w = SampleCosineHemisphere(vec2(u2[2 * i], u2[2 * i + 1])); w3[3 * i] = w.x; w3[3 * i + 1] = w.y; w3[3 * i + 2] = w.z; pdf[i] = CosineHemispherePDF(w.z); }
}
// SampledWavelengths::TerminateSecondary (spectrum.h:302-310) applied to SampleVisible(u): the pdfs afterwards
void orc_terminate_secondary(float u, float* pdf8) {
    SampledWavelengths w = SampledWavelengths::SampleVisible(u);
    w.TerminateSecondary();
    for (int i = 0; i < 8; ++i) pdf8[i] = w.pdf[i];
}
// SampleLinear(u, a, b) (RayTracer/Sampling.h:205-211): the two deterministic halves of SampleTent -- (0,1) and (1,0) -- given its coin
float orc_sample_linear(float u, float a, float b) { return SampleLinear(u, a, b); }
void orc_concentric_disk(float u0, float u1, float* out2) { vec2 d = SampleUniformDiskConcentric(vec2(u0, u1)); out2[0] = d.x; out2[1] = d.y; }
float orc_gamma(int n) { return gamma_n(n); }
float orc_difference_of_products(float a, float b, float c, float d) { return DifferenceOfProducts(a, b, c, d); }

// spectral tables and colour constants the product must reproduce bit for bit
// which: 0 X, 1 Y, 2 Z (dense 471), 3 sRGB illuminant D65 dense (471)
void orc_dense_table(int which, float* out471) {
    const auto& T = SpectraTables::get();
    const DenselySampledSpectrum* d = which == 0 ? T.X.get() : which == 1 ? T.Y.get() : which == 2 ? T.Z.get() : &RGBColorSpace::sRGB().illuminant;
    for (int i = 0; i < 471; ++i) out471[i] = d->values[i];
}
// normalised illuminant knots: which 0 A, 1 D50, 2 D65, 3 F1, 4 F2, 5 F11; returns knot count
int orc_illuminant_knots(int which, float* lambdas, float* values, int cap) {
    const auto& T = SpectraTables::get();
    const PiecewiseLinearSpectrum* p = which == 0 ? T.illumA.get() : which == 1 ? T.illumD50.get() : which == 2 ? T.illumD65.get()
                                     : which == 3 ? T.illumF1.get() : which == 4 ? T.illumF2.get() : T.illumF11.get();
    int n = (int)p->lambdas.size();
    for (int i = 0; i < n && i < cap; ++i) { lambdas[i] = p->lambdas[i]; values[i] = p->values[i]; }
    return n;
}
// out: XYZFromSensorRGB (9, column-major), RGBFromXYZ (9), XYZFromRGB (9), white xy (2)
void orc_color_constants(float* sensor9, float* rgbfromxyz9, float* xyzfromrgb9, float* white2) {
    const RGBColorSpace& cs = RGBColorSpace::sRGB();
    PixelSensor sensor(cs, SpectraTables::get().illumD65.get(), 1.0f / CIE_Y_integral);
    std::memcpy(sensor9, sensor.XYZFromSensorRGB.c, 36);
    std::memcpy(rgbfromxyz9, cs.RGBFromXYZ.c, 36);
    std::memcpy(xyzfromrgb9, cs.XYZFromRGB.c, 36);
    white2[0] = cs.w.x; white2[1] = cs.w.y;
}
// RGBToSpectrumTable(zNodes, coeffs): scale[64], data[3][64][64][64][3] as the reference's Init hands them over (color.cpp:164)
void orc_set_rgb_table(const float* scale64, const float* data) {
    auto& t = RGBToSpectrumTable::sRGB();
    if (!scale64) { t.zNodes.clear(); t.coeffs.clear(); return; }
    t.zNodes.assign(scale64, scale64 + 64);
    t.coeffs.assign(data, data + (size_t)3 * 64 * 64 * 64 * 3);
}
// RGBColorSpace::ToRGBCoeffs via RGBAlbedoSpectrum: rgb -> (c0, c1, c2); -1 if non-grey without a table
int orc_rgb_coeffs(const float* rgb3, float* c3) {
    RGBAlbedoSpectrum a;
    if (!MakeRGBAlbedo(rgb3[0], rgb3[1], rgb3[2], &a)) return -1;
    c3[0] = a.rsp.c0; c3[1] = a.rsp.c1; c3[2] = a.rsp.c2;
    return 0;
}
// Measured PixelSensor (pixelsensor.h:37-68) from response curves and a sensor illuminant sampled at 360..830 nm (what the
// constructor's DenselySampledSpectrum members and 1 nm integrals see).  r471 == NULL restores the XYZ sensor.  matrix9_out:
// XYZFromSensorRGB, column-major.
void orc_set_sensor(const float* r471, const float* g471, const float* b471, const float* illum471, float imaging_ratio, float* matrix9_out) {
    if (!r471) { g_sensor_override.reset(); return; }
    DenseFromArray r(r471), g(g471), b(b471), il(illum471);
    g_sensor_override = std::make_unique<PixelSensor>(&r, &g, &b, RGBColorSpace::sRGB(), &il, imaging_ratio);
    if (matrix9_out) std::memcpy(matrix9_out, g_sensor_override->XYZFromSensorRGB.c, 36);
}
float orc_sigmoid_eval(float c0, float c1, float c2, float lambda) { return RGBSigmoidPolynomial{c0, c1, c2}(lambda); }
int orc_grey_sigmoid(float g, float* c3) { RGBSigmoidPolynomial p; if (!GreyToSigmoid(g, g, g, &p)) return -1; c3[0] = p.c0; c3[1] = p.c1; c3[2] = p.c2; return 0; }

// camera matrices (Cameras.h:77-142,248-311).  kind 0 perspective, 1 orthographic
void orc_camera_matrices(int kind, float near_, float far_, float sw, float sh, float fov, const float* pos, const float* look,
                         const float* right, const float* up, float resx, float resy, float* r2c16, float* c2w16) {
    vec3 p(pos[0], pos[1], pos[2]), l(look[0], look[1], look[2]), r(right[0], right[1], right[2]), u(up[0], up[1], up[2]);
    std::unique_ptr<CameraBase> cam;
    if (kind == 0) cam = std::make_unique<PerspectiveCamera>(near_, far_, sw, sh, fov, p, l, r, u, vec2(resx, resy));
    else if (kind == 1) cam = std::make_unique<OrthographicCamera>(near_, far_, sw, sh, p, l, r, u, vec2(resx, resy));
    else {
        cam = std::make_unique<PinholeCamera>(0.25f, vec3(sw, sh, far_), p, l, r, u, vec2(resx, resy));
        std::memcpy(r2c16, cam->M_RastertoScreen.c, 64);
        std::memcpy(c2w16, cam->M_CameratoWorld.c, 64);
        return;
    }
    std::memcpy(r2c16, cam->M_RastertoCamera.c, 64);
    std::memcpy(c2w16, cam->M_CameratoWorld.c, 64);
}
// Shape transform convention (Shapes.h:175-182): rigid -> ObjectToRender, RenderToObject
void orc_shape_matrices(const float* rigid16, float* o2r16, float* r2o16) {
    struct S : Shape {
        using Shape::Shape;
        Bounds3 Bounds() const override { return {}; }
        std::optional<LocalSurfaceInfo> Intersect(const Ray&, float) const override { return {}; }
        bool IntersectP(const Ray&, float) const override { return false; }
        float Area() const override { return 0; }
    } s("m", mat4::from_ptr(rigid16));
    std::memcpy(o2r16, s.ObjectToRender.c, 64);
    std::memcpy(r2o16, s.RenderToObject.c, 64);
}

// ---------------------------------------------------------------- scene --------------------------
void* orc_scene_create() {
    SpectraTables::get();
    RGBColorSpace::sRGB();
    auto* s = new OScene;
    s->model_name = "oracle_scene_" + std::to_string(g_scene_counter++);
    return s;
}
void orc_scene_destroy(void* h) {
    auto* s = (OScene*)h;
    MeshCache::modelCache().erase(s->model_name);
    delete s;
}
// meshes are concatenated: positions/normals 3 floats per vertex (normals may be null), indices per mesh are local
int orc_scene_set_model(void* h, int n_meshes, const float* positions, const float* normals, const uint32_t* nverts,
                        const uint32_t* indices, const uint32_t* ntris, const float* rigid16, int precomputed_world,
                        int cull_backface, const float* look_dir, const float* texcoords, const float* tangents, const float* bitangents) {
    auto* s = (OScene*)h;
    MeshCache::Model model;
    model.mesh_name = s->model_name;
    size_t vo = 0, io = 0;
    for (int m = 0; m < n_meshes; ++m) {
        MeshCache::Mesh mesh;
        for (uint32_t v = 0; v < nverts[m]; ++v) {
            mesh.positions.push_back(vec3(positions[3 * (vo + v)], positions[3 * (vo + v) + 1], positions[3 * (vo + v) + 2]));
            if (normals) mesh.normals.push_back(vec3(normals[3 * (vo + v)], normals[3 * (vo + v) + 1], normals[3 * (vo + v) + 2]));
            if (texcoords) mesh.texcoords.push_back(vec2(texcoords[2 * (vo + v)], texcoords[2 * (vo + v) + 1]));
            if (tangents) mesh.tangents.push_back(vec3(tangents[3 * (vo + v)], tangents[3 * (vo + v) + 1], tangents[3 * (vo + v) + 2]));
            if (bitangents) mesh.bitangents.push_back(vec3(bitangents[3 * (vo + v)], bitangents[3 * (vo + v) + 1], bitangents[3 * (vo + v) + 2]));
        }
        mesh.indices.assign(indices + io, indices + io + 3 * (size_t)ntris[m]);
        vo += nverts[m]; io += 3 * (size_t)ntris[m];
        model.meshes.push_back(std::move(mesh));
    }
    MeshCache::modelCache()[s->model_name] = std::move(model);
    Triangle::vertex_available avail;
    avail.texcoords = texcoords != nullptr; avail.tangents = tangents != nullptr; avail.bitangents = bitangents != nullptr;
    avail.normals = normals != nullptr;
    avail.precomputed_worldtransform = precomputed_world != 0;
    s->model = std::make_unique<TriModel>("model", mat4::from_ptr(rigid16), s->model_name, cull_backface != 0, precomputed_world != 0, avail);
    if (cull_backface && normals) s->model->ComputeBackFace(vec3(look_dir[0], look_dir[1], look_dir[2]), true);
    s->scene.mesh_material.assign(n_meshes, 0);
    return 0;
}
int orc_scene_build_octree(void* h) {
    auto* s = (OScene*)h;
    s->oct = std::make_unique<Octtree_Model>(*s->model);
    s->oct->CreateOcttree();
    s->scene.oct = s->oct.get();
    return s->oct->getTreeSize();
}
void orc_octree_stats(void* h, int* ints7, float* avg_leaf, long long* refs) {
    auto st = ((OScene*)h)->oct->GetStats();
    ints7[0] = st.nodes; ints7[1] = st.real_nodes; ints7[2] = st.leaves; ints7[3] = st.empty_leaves; ints7[4] = st.max_leaf; ints7[5] = st.depth; ints7[6] = 0;
    *avg_leaf = st.avg_leaf; *refs = st.refs;
}
// reference-order node dump: bounds (6/node), leaf flag, child ids (8/node, -1 for leaves), per-node list offset/count
long long orc_octree_dump(void* h, float* bounds6, int32_t* leaf, int32_t* child8, long long* list_off, int32_t* list_pairs, long long cap_pairs) {
    auto* s = (OScene*)h;
    long long off = 0;
    for (size_t i = 0; i < s->oct->octtree.size(); ++i) {
        const auto& n = s->oct->octtree[i];
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
    if (list_off) list_off[s->oct->octtree.size()] = off;
    return off;
}
int orc_model_backfacing(void* h, int mesh, unsigned char* out) {
    auto* s = (OScene*)h;
    if (s->model->back_facing.empty()) return 0;
    const auto& v = s->model->back_facing[mesh];
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i] ? 1 : 0;
    return (int)v.size();
}
void orc_model_bounds(void* h, float* out6) {
    Bounds3 b = ((OScene*)h)->model->Bounds();
    out6[0] = b.pmin.x; out6[1] = b.pmin.y; out6[2] = b.pmin.z; out6[3] = b.pmax.x; out6[4] = b.pmax.y; out6[5] = b.pmax.z;
}

// kind: 0 sphere(r,zmin,zmax,phimax) 1 cylinder(r,zmin,zmax,phimax) 2 disk(h,inner,outer,phimax) 3 trianglesimple(p1,p2,p3)
int orc_scene_add_shape(void* h, int kind, const float* rigid16, const float* params, int material) {
    auto* s = (OScene*)h;
    mat4 M = mat4::from_ptr(rigid16);
    std::unique_ptr<Shape> sh;
    if (kind == 0) sh = std::make_unique<Sphere>("s", M, params[0], params[1], params[2], params[3]);
    else if (kind == 1) sh = std::make_unique<Cylinder>("c", M, params[0], params[1], params[2], params[3]);
    else if (kind == 2) sh = std::make_unique<Disk>("d", M, params[0], params[1], params[2], params[3]);
    else if (kind == 3) sh = std::make_unique<TriangleSimple>("t", M, vec3(params[0], params[1], params[2]), vec3(params[3], params[4], params[5]), vec3(params[6], params[7], params[8]));
    else return -1;
    s->scene.shapes.push_back(sh.get());
    s->scene.shape_material.push_back(material);
    s->shapes.push_back(std::move(sh));
    return (int)s->shapes.size() - 1;
}
// spectra: kind 0 constant(c); 1 piecewise from interleaved (lambda,value) pairs; 2 named table; 3 swatch i;
// 4 normalised illuminant (which); 5 grey RGB albedo(g); 6 grey RGB illuminant(g) (x D65 dense);
// 7 / 8 / 9 RGBAlbedo / RGBIlluminant / RGBUnbounded of rgb = interleaved[0..2] (needs orc_set_rgb_table unless grey)
int orc_scene_add_spectrum(void* h, int kind, float c, const float* interleaved, int n, const char* name, int normalize) {
    auto* s = (OScene*)h;
    std::unique_ptr<Spectrum> sp;
    const auto& T = SpectraTables::get();
    if (kind == 0) sp = std::make_unique<ConstantSpectrum>(c);
    else if (kind == 1) sp.reset(PiecewiseLinearSpectrum::FromInterleaved(interleaved, n, normalize != 0));
    else if (kind == 2) { int cnt; const float* t = named_table(name, &cnt); if (!t) return -1; sp.reset(PiecewiseLinearSpectrum::FromInterleaved(t, cnt, normalize != 0)); }
    else if (kind == 3) { int cnt; const float* t = swatch_table(n, &cnt); sp.reset(PiecewiseLinearSpectrum::FromInterleaved(t, cnt, false)); }
    else if (kind == 4) {
        const PiecewiseLinearSpectrum* p = n == 0 ? T.illumA.get() : n == 1 ? T.illumD50.get() : n == 2 ? T.illumD65.get() : n == 3 ? T.illumF1.get() : n == 4 ? T.illumF2.get() : T.illumF11.get();
        sp = std::make_unique<PiecewiseLinearSpectrum>(*p);
    } else if (kind == 5) { auto a = std::make_unique<RGBAlbedoSpectrum>(); if (!MakeRGBAlbedo(c, c, c, a.get())) return -1; sp = std::move(a); }
    else if (kind == 6) { auto a = std::make_unique<RGBIlluminantSpectrum>(); if (!MakeRGBIlluminant(c, c, c, a.get())) return -1; sp = std::move(a); }
    else if (kind == 7) { auto a = std::make_unique<RGBAlbedoSpectrum>(); if (!MakeRGBAlbedo(interleaved[0], interleaved[1], interleaved[2], a.get())) return -1; sp = std::move(a); }
    else if (kind == 8) { auto a = std::make_unique<RGBIlluminantSpectrum>(); if (!MakeRGBIlluminant(interleaved[0], interleaved[1], interleaved[2], a.get())) return -1; sp = std::move(a); }
    else if (kind == 9) { auto a = std::make_unique<RGBUnboundedSpectrum>(); if (!MakeRGBUnbounded(interleaved[0], interleaved[1], interleaved[2], a.get())) return -1; sp = std::move(a); }
    else return -1;
    s->scene.spectra.push_back(std::move(sp));
    return (int)s->scene.spectra.size() - 1;
}
void orc_spectrum_sample(void* h, int id, const float* lambda8, float* out8) {
    auto* s = (OScene*)h;
    for (int i = 0; i < 8; ++i) out8[i] = s->scene.spectra[id]->Query(lambda8[i]);
}
int orc_scene_add_material(void* h, int type, int refl, int eta, int k, int emit, float emit_scale, int two_sided, int eta_constant) {
    auto* s = (OScene*)h;
    Material m;
    m.type = type; m.refl = refl; m.eta = eta; m.k = k; m.emit = emit; m.emit_scale = emit_scale; m.two_sided = two_sided; m.eta_constant = eta_constant != 0;
    s->scene.materials.push_back(m);
    return (int)s->scene.materials.size() - 1;
}
// Lights.h:5-8: kind 0 point light (v = position), 1 sun (v = direction towards the light, normalised here)
int orc_scene_add_light(void* h, int kind, const float* v3, int spectrum, float scale) {
    auto* s = (OScene*)h;
    DeltaLight dl;
    dl.kind = kind; dl.v = vec3(v3[0], v3[1], v3[2]); dl.spectrum = spectrum; dl.scale = scale;
    if (kind == 1) dl.v = normalize(dl.v);
    s->scene.delta_lights.push_back(dl);
    return (int)s->scene.delta_lights.size() - 1;
}
void orc_scene_set_mesh_materials(void* h, const int* ids, int n) { ((OScene*)h)->scene.mesh_material.assign(ids, ids + n); }
int orc_scene_light_count(void* h) { auto* s = (OScene*)h; s->scene.BuildLights(); return (int)s->scene.lights.size(); }
void orc_scene_light_cdf(void* h, float* cdf, int32_t* mesh_tri) {
    auto* s = (OScene*)h;
    for (size_t i = 0; i < s->scene.lights.size(); ++i) { cdf[i] = s->scene.light_cdf[i]; mesh_tri[2 * i] = s->scene.lights[i].mesh_id; mesh_tri[2 * i + 1] = s->scene.lights[i].tri_id; }
}

// ---------------------------------------------------------------- traversal probes ----------------
static void run_parallel(int n, int nthreads, const std::function<void(int, int, int)>& fn) {
    nthreads = std::max(1, std::min(nthreads, n > 0 ? n : 1));
    std::vector<std::thread> pool;
    int per = n / nthreads, b = 0;
    for (int t = 0; t < nthreads; ++t) { int e = (t == nthreads - 1) ? n : b + per; pool.emplace_back(fn, t, b, e); b = e; }
    for (auto& th : pool) th.join();
}
// mode 0: octree BFS (Octtree_Model::Traverse); 1: brute force (TriModel::BasicIntersect); 2: any-hit with tmax[i]
// rays: 6 floats (o, d).  counters5: rays, nodes, tris, leaves, max_queue (may be null)
void orc_trace(void* h, int mode, const float* rays, int n, const float* tmax, int32_t* mesh_id, int32_t* tri_id, float* t, float* bary3,
               int nthreads, unsigned long long* counters5) {
    auto* s = (OScene*)h;
    std::vector<TraverseCounters> tcs(std::max(1, nthreads));
    run_parallel(n, nthreads, [&](int tid, int b, int e) {
        tl_counters = counters5 ? &tcs[tid] : nullptr;
        for (int i = b; i < e; ++i) {
            Ray ray(vec3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), vec3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));
            if (mode == 0) {
                auto rec = s->oct->TraverseClosest(ray);
                mesh_id[i] = rec.found ? rec.info.mesh_id : -1; tri_id[i] = rec.found ? rec.info.tri_id : -1;
                if (t) t[i] = rec.found ? rec.isect.t : 0;
                if (bary3) { bary3[3 * i] = rec.isect.b0; bary3[3 * i + 1] = rec.isect.b1; bary3[3 * i + 2] = rec.isect.b2; }
            } else if (mode == 1) {
                auto r = s->model->BasicIntersect(ray);
                mesh_id[i] = r ? r->mesh_id : -1; tri_id[i] = r ? r->tri_id : -1;
                if (t) t[i] = r ? r->tri_isect.t : 0;
                if (bary3 && r) { bary3[3 * i] = r->tri_isect.b0; bary3[3 * i + 1] = r->tri_isect.b1; bary3[3 * i + 2] = r->tri_isect.b2; }
            } else {
                mesh_id[i] = s->scene.Occluded(ray, tmax[i]) ? 1 : 0;
            }
        }
        tl_counters = nullptr;
    });
    if (counters5) {
        TraverseCounters tot;
        for (auto& c : tcs) tot.add(c);
        counters5[0] = tot.rays; counters5[1] = tot.nodes_visited; counters5[2] = tot.tris_tested; counters5[3] = tot.leaves_visited; counters5[4] = tot.max_queue;
    }
}
// Octtree_Model::Traverse incl. CalculateLocalSurface: n (3), hitp (3), u, v per ray; found flag
void orc_traverse_surface(void* h, const float* rays, int n, int32_t* found, float* nrm3, float* hitp3, float* uv2) {
    auto* s = (OScene*)h;
    for (int i = 0; i < n; ++i) {
        Ray ray(vec3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), vec3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));
        auto r = s->oct->Traverse(ray);
        found[i] = r.has_value();
        if (r) { nrm3[3 * i] = r->n.x; nrm3[3 * i + 1] = r->n.y; nrm3[3 * i + 2] = r->n.z; hitp3[3 * i] = r->hitp.x; hitp3[3 * i + 1] = r->hitp.y; hitp3[3 * i + 2] = r->hitp.z; uv2[2 * i] = r->u; uv2[2 * i + 1] = r->v; }
    }
}
static void put_info17(const LocalSurfaceInfo& r, float* w) {
    w[0] = r.hitp.x; w[1] = r.hitp.y; w[2] = r.hitp.z; w[3] = r.u; w[4] = r.v;
    w[5] = r.du.x; w[6] = r.du.y; w[7] = r.du.z; w[8] = r.dv.x; w[9] = r.dv.y; w[10] = r.dv.z;
    w[11] = r.n.x; w[12] = r.n.y; w[13] = r.n.z; w[14] = r.wo.x; w[15] = r.wo.y; w[16] = r.wo.z;
}
// Octtree_Model::Traverse, whole LocalSurfaceInfo record (Shapes.h:144-170) minus the never-assigned tHit: 17 floats per ray
void orc_traverse_local_surface(void* h, const float* rays, int n, int32_t* found, float* info17) {
    auto* s = (OScene*)h;
    for (int i = 0; i < n; ++i) {
        Ray ray(vec3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), vec3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));
        auto r = s->oct->Traverse(ray);
        found[i] = r.has_value();
        if (r) put_info17(*r, info17 + 17 * (size_t)i);
    }
}
// Triangle(mesh, tri).CalculateLocalSurface for given barycentrics and (normalised) ray direction, bypassing BasicIntersect
void orc_local_surface_of(void* h, const int32_t* mesh_id, const int32_t* tri_id, const float* bary3, const float* rayd3, int n, float* info17) {
    auto* s = (OScene*)h;
    for (int i = 0; i < n; ++i) {
        const Triangle& tri = s->model->triangles[mesh_id[i]][tri_id[i]];
        Triangle::TriangleIntersect is;
        is.b0 = bary3[3 * i]; is.b1 = bary3[3 * i + 1]; is.b2 = bary3[3 * i + 2]; is.t = 0; is.rayd = vec3(rayd3[3 * i], rayd3[3 * i + 1], rayd3[3 * i + 2]);
        auto r = tri.CalculateLocalSurface(is);
        if (r) put_info17(*r, info17 + 17 * (size_t)i);
    }
}
// Scene::Closest (mesh + analytic shapes): kind (-1 miss, 0 tri, 1 shape), ids, t, p, ns_ff, ng_ff, backside
void orc_scene_closest(void* h, const float* rays, int n, int32_t* kind, int32_t* id0, int32_t* id1, float* t, float* p3, float* ns3, float* ng3, int32_t* backside) {
    auto* s = (OScene*)h;
    for (int i = 0; i < n; ++i) {
        Ray ray(vec3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), vec3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));
        SurfaceHit sh = s->scene.Closest(ray);
        kind[i] = sh.found ? sh.kind : -1;
        id0[i] = sh.kind == 0 ? sh.mesh_id : sh.shape_id; id1[i] = sh.tri_id;
        t[i] = sh.t;
        for (int k = 0; k < 3; ++k) { p3[3 * i + k] = sh.p[k]; ns3[3 * i + k] = sh.ns_ff[k]; ng3[3 * i + k] = sh.ng_ff[k]; }
        backside[i] = sh.backside;
    }
}
// Shape::Area() (Shapes.h:198; per shape :234,:455,:642,:779) and Triangle::Area() (:949-961) of triangle (mesh, tri)
// Shape::Intersect, the rest of the record: du (3), dv (3), wo (3) per ray (zero where nothing was hit)
void orc_shape_frame(void* h, int shape, const float* rays, int n, float tmax, float* du_dv_wo9) {
    auto* s = (OScene*)h;
    for (int i = 0; i < n; ++i) {
        Ray ray(vec3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), vec3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));
        auto r = s->shapes[shape]->Intersect(ray, tmax);
        if (r) for (int k = 0; k < 3; ++k) { du_dv_wo9[9 * i + k] = r->du[k]; du_dv_wo9[9 * i + 3 + k] = r->dv[k]; du_dv_wo9[9 * i + 6 + k] = r->wo[k]; }
    }
}
float orc_shape_area(void* h, int shape) { return ((OScene*)h)->shapes[shape]->Area(); }
void orc_triangle_area(void* h, const int32_t* mesh_id, const int32_t* tri_id, int n, float* out) {
    auto* s = (OScene*)h;
    for (int i = 0; i < n; ++i) out[i] = s->model->triangles[mesh_id[i]][tri_id[i]].Area();
}
// Moller::triBoxOverlap (AABB_triangle_Moller.h:229-474) on n (center, half size, triangle) triples
void orc_tri_box_overlap(const float* center3, const float* half3, const float* tri9, int n, int32_t* out) {
    for (int i = 0; i < n; ++i) {
        vec3 tv[3];
        for (int a = 0; a < 3; ++a) tv[a] = vec3(tri9[9 * i + 3 * a], tri9[9 * i + 3 * a + 1], tri9[9 * i + 3 * a + 2]);
        out[i] = moller::triBoxOverlap(vec3(center3[3 * i], center3[3 * i + 1], center3[3 * i + 2]), vec3(half3[3 * i], half3[3 * i + 1], half3[3 * i + 2]), tv);
    }
}
// single analytic shape probes (Shape::Intersect): found, t, hitp, n, u, v
void orc_shape_intersect(void* h, int shape, const float* rays, int n, float tmax, int32_t* found, float* t, float* hitp3, float* nrm3, float* uv2) {
    auto* s = (OScene*)h;
    for (int i = 0; i < n; ++i) {
        Ray ray(vec3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), vec3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));
        auto r = s->shapes[shape]->Intersect(ray, tmax);
        found[i] = r.has_value();
        if (r) { t[i] = r->tHit; for (int k = 0; k < 3; ++k) { hitp3[3 * i + k] = r->hitp[k]; nrm3[3 * i + k] = r->n[k]; } uv2[2 * i] = r->u; uv2[2 * i + 1] = r->v; }
    }
}

// ---------------------------------------------------------------- rendering -----------------------
struct orc_render_params {
    int width, height;
    float r2c[16], c2w[16];
    float lens_radius, focal_distance;
    int camera_kind;               // 0 perspective, 1 orthographic
    int sampler_kind, xs, ys, jitter, seed;   // 0 independent (spp = xs*ys), 1 stratified
    int filter_kind;               // 0 box, 1 triangle, 2 gaussian (filter_sigma below; 0 = the class default 0.5)
    float filter_rx, filter_ry;
    int mode, max_depth, rr_depth; // IntegratorConfig
    float ray_eps, shadow_eps;
    float albedo[3];
    int spp_begin, spp_end;
    int nthreads, pixel_stride;
    int faithful_overheads;
    float filter_sigma;
    int light_strategy;
};
struct OrthoMatrixCamera : CameraBase {
    OrthoMatrixCamera(const mat4& r2c, const mat4& c2w) : CameraBase(1, 1, vec3(0, 0, 0), vec3(0, 0, 1), vec3(1, 0, 0), vec3(0, 1, 0), vec2(1, 1)) { M_RastertoCamera = r2c; M_CameratoWorld = c2w; }
    Ray generateRay(vec2 pixel, Sampler*) override {
        vec4 c = mul(M_RastertoCamera, vec4(pixel.x, pixel.y, 0, 1));
        Ray ray(xyz(c), vec3(0, 0, 1));
        ray.Transform(M_CameratoWorld);
        return ray;
    }
};
// PinholeCamera (Cameras.h:313-359) given by M_RastertoScreen (in the r2c slot), M_CameratoWorld and the box depth
struct PinholeMatrixCamera : CameraBase {
    float box_z;
    PinholeMatrixCamera(const mat4& r2s, const mat4& c2w, float bz) : CameraBase(1, 1, vec3(0, 0, 0), vec3(0, 0, 1), vec3(1, 0, 0), vec3(0, 1, 0), vec2(1, 1)), box_z(bz) { M_RastertoScreen = r2s; M_CameratoWorld = c2w; }
    Ray generateRay(vec2 pixel, Sampler*) override {
        vec3 sensor_pos = xyz(mul(M_RastertoScreen, vec4(pixel.x, pixel.y, 0, 1)));
        float hole_radius = 0.25f;
        vec3 pinhole(0.0f * hole_radius * std::cos(radians(0.0f)), 0.0f * hole_radius * std::sin(radians(0.0f)), box_z);
        Ray ray(sensor_pos, normalize(pinhole - sensor_pos));
        ray.Transform(M_CameratoWorld);
        return ray;
    }
};
struct RenderCtx {
    std::unique_ptr<CameraBase> cam;
    std::unique_ptr<Sampler> sampler;
    std::unique_ptr<Filter> filter;
    std::unique_ptr<PixelSensor> sensor;
    Film film;
    Renderer r;
};
static void setup(OScene* s, const orc_render_params* p, RenderCtx& c) {
    if (p->camera_kind == 0) c.cam = std::make_unique<MatrixPerspectiveCamera>(mat4::from_ptr(p->r2c), mat4::from_ptr(p->c2w), p->lens_radius, p->focal_distance);
    else if (p->camera_kind == 1) c.cam = std::make_unique<OrthoMatrixCamera>(mat4::from_ptr(p->r2c), mat4::from_ptr(p->c2w));
    else c.cam = std::make_unique<PinholeMatrixCamera>(mat4::from_ptr(p->r2c), mat4::from_ptr(p->c2w), p->focal_distance);
    c.sampler = make_sampler(p->sampler_kind, p->xs, p->ys, p->jitter, p->seed);
    if (p->filter_kind == 0) c.filter = std::make_unique<BoxFilter>(vec2(p->filter_rx, p->filter_ry));
    else if (p->filter_kind == 2) c.filter = std::make_unique<GaussianFilter>(vec2(p->filter_rx, p->filter_ry), p->filter_sigma > 0 ? p->filter_sigma : 0.5f);
    else c.filter = std::make_unique<TriangleFilter>(vec2(p->filter_rx, p->filter_ry));
    c.sensor = make_sensor();
    c.film.image_res = ivec2(p->width, p->height);
    c.film.film_dim = ivec2(p->width, p->height);
    c.film.pixels.assign((size_t)p->width * p->height, pixel());
    c.film.filter = c.filter.get();
    c.film.pixel_sensor = c.sensor.get();
    c.r.scene = &s->scene; c.r.camera = c.cam.get(); c.r.film = &c.film;
    c.r.cfg.mode = p->mode; c.r.cfg.max_depth = p->max_depth; c.r.cfg.rr_depth = p->rr_depth;
    c.r.cfg.ray_eps = p->ray_eps; c.r.cfg.shadow_eps = p->shadow_eps; c.r.cfg.light_strategy = p->light_strategy;
    for (int i = 0; i < 3; ++i) c.r.cfg.albedo_rgb[i] = p->albedo[i];
    c.r.Prepare();
}
// film_io: width*height*4 floats (rgbsum, weightsum); accumulated INTO (pass zeros for a fresh render).
// counters9: rays, nodes, tris, leaves, max_queue, paths, closest_rays, shadow_rays, depth_sum
double orc_render(void* h, const orc_render_params* p, float* film_io, unsigned long long* counters9) {
    auto* s = (OScene*)h;
    RenderCtx c;
    setup(s, p, c);
    size_t np = c.film.pixels.size();
    for (size_t i = 0; i < np; ++i) { c.film.pixels[i].rgbsum = vec3(film_io[4 * i], film_io[4 * i + 1], film_io[4 * i + 2]); c.film.pixels[i].weightsum = film_io[4 * i + 3]; }
    TraverseCounters tc; PathCounters pc;
    g_faithful_overheads = p->faithful_overheads != 0;
    double secs = c.r.RenderThreaded(*c.sampler, p->spp_begin, p->spp_end, p->nthreads, p->pixel_stride, counters9 ? &tc : nullptr, counters9 ? &pc : nullptr);
    g_faithful_overheads = false;
    for (size_t i = 0; i < np; ++i) { film_io[4 * i] = c.film.pixels[i].rgbsum.x; film_io[4 * i + 1] = c.film.pixels[i].rgbsum.y; film_io[4 * i + 2] = c.film.pixels[i].rgbsum.z; film_io[4 * i + 3] = c.film.pixels[i].weightsum; }
    if (counters9) {
        counters9[0] = tc.rays; counters9[1] = tc.nodes_visited; counters9[2] = tc.tris_tested; counters9[3] = tc.leaves_visited; counters9[4] = tc.max_queue;
        counters9[5] = pc.paths; counters9[6] = pc.closest_rays; counters9[7] = pc.shadow_rays; counters9[8] = pc.depth_sum;
    }
    return secs;
}
// per-sample probe for parity: for each (pixel_id, index): ray (6), lambda (8), pdf (8), L (8), clamped sensor rgb (3), weight
void orc_eval_samples(void* h, const orc_render_params* p, const int32_t* pixel_ids, const int32_t* indices, int n, float* ray6, float* lambda8,
                      float* pdf8, float* L8, float* rgb3, float* weight) {
    auto* s = (OScene*)h;
    RenderCtx c;
    setup(s, p, c);
    for (int i = 0; i < n; ++i) {
        Renderer::SampleDebug d;
        c.r.evaluate_pixel(pixel_ids[i], indices[i], c.sampler.get(), &d, nullptr);
        for (int k = 0; k < 3; ++k) { ray6[6 * i + k] = d.ray.o[k]; ray6[6 * i + 3 + k] = d.ray.d[k]; rgb3[3 * i + k] = d.rgb[k]; }
        for (int k = 0; k < 8; ++k) { lambda8[8 * i + k] = d.lambdas.lambda[k]; pdf8[8 * i + k] = d.lambdas.pdf[k]; L8[8 * i + k] = d.L[k]; }
        weight[i] = d.weight;
    }
}
// film resolve (RayTracerTestApp.h:425-452): film4 -> rgb8 and/or float rgb
void orc_resolve(const float* film4, int npix, unsigned char* rgb8, float* rgbf) {
    std::unique_ptr<PixelSensor> sp = make_sensor();
    PixelSensor& sensor = *sp;
    Film f;
    f.pixels.resize(npix);
    for (int i = 0; i < npix; ++i) { f.pixels[i].rgbsum = vec3(film4[4 * i], film4[4 * i + 1], film4[4 * i + 2]); f.pixels[i].weightsum = film4[4 * i + 3]; }
    f.pixel_sensor = &sensor;
    ResolveFilm(f, RGBColorSpace::sRGB().RGBFromXYZ, rgb8, rgbf);
}
int orc_hardware_threads() { return (int)std::max(1u, std::thread::hardware_concurrency()); }

}  // extern "C"
