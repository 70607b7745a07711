// ORACLE (test infrastructure, never shipped, never on the product path).
//
// Cameras (RayTracer/Cameras.h), Film (RayTracer/Film.h) and the renderer that lives in
// Applications/RayTracerTestApp.h:218-452 (Li, evaluate_pixel, thread pool, resolve) -- "Tier A",
// restated from reference code -- plus the "Tier B" path integrator the north-star asks for and the
// reference only sketches in comments (Integrator.h:1-13, Shading.h:1-20, Lights.h:1-9).  Tier B has
// no reference implementation: this file IS its definition and the CUDA path must match it.
#pragma once
#include <atomic>
#include <thread>

#include "oracle_shapes.h"

namespace orc {

// ---- Cameras.h:77-211 ------------------------------------------------------------------------
struct CameraBase {
    float lensRadius, focalDistance;
    vec3 world_pos, look_direction, right_direction, up_direction, worldup_direction;
    vec2 sensor_dimensions, image_resolution;
    mat4 M_ScreentoRaster, M_RastertoScreen, M_CameratoWorld, M_WorldtoCamera, M_CameratoScreen, M_RastertoCamera;
    CameraBase(float sw, float sh, vec3 pos, vec3 look, vec3 right, vec3 worldup, vec2 res, float lensR = 0, float focalD = 0)
        : lensRadius(lensR), focalDistance(focalD), world_pos(pos), look_direction(look), right_direction(right),
          worldup_direction(worldup), sensor_dimensions(sw, sh), image_resolution(res) {
        up_direction = cross(look_direction, right_direction);
        mat4 I;
        mat4 S2N = mul(scale(I, vec3(1.0f / sensor_dimensions.x, 1.0f / sensor_dimensions.y, 1)),
                       translate(I, vec3(sensor_dimensions.x / 2.0f, sensor_dimensions.y / 2.0f, 0)));
        mat4 N2R = mul(scale(I, vec3(image_resolution.x, -image_resolution.y, 1)), translate(I, vec3(0, -1, 0)));
        M_ScreentoRaster = mul(N2R, S2N);
        M_RastertoScreen = inverse(M_ScreentoRaster);
        calculateWorldCameraMatrices();
    }
    virtual ~CameraBase() = default;
    void calculateWorldCameraMatrices() {          // :130-142
        vec3 dir = normalize(look_direction);
        right_direction = normalize(cross(worldup_direction, dir));
        up_direction = cross(dir, right_direction);
        M_CameratoWorld = mat4::from_cols(vec4(right_direction, 0), vec4(up_direction, 0), vec4(dir, 0), vec4(world_pos, 1));
        M_WorldtoCamera = inverse(M_CameratoWorld);
    }
    void setWorldPos(vec3 p) { world_pos = p; calculateWorldCameraMatrices(); }
    void SetlensRadius(float r) { lensRadius = r; }
    void SetfocalDistance(float d) { focalDistance = d; }
    virtual Ray generateRay(vec2 pixel, Sampler* sampler) = 0;
};
struct OrthographicCamera : CameraBase {           // :213-245
    OrthographicCamera(float N, float F, float sw, float sh, vec3 pos, vec3 look, vec3 right, vec3 up, vec2 res)
        : CameraBase(sw, sh, pos, look, right, up, res) {
        mat4 I;
        M_CameratoScreen = mul(scale(I, vec3(1, 1, (float)(1.0 / (F - N)))), translate(I, vec3(0, 0, -N)));
        M_RastertoCamera = mul(inverse(M_CameratoScreen), M_RastertoScreen);
    }
    Ray generateRay(vec2 pixel, Sampler*) override {
        vec4 c = mul(M_RastertoCamera, vec4(pixel.x, pixel.y, 0, 1));
        Ray ray(xyz(c), vec3(0, 0, 1));
        ray.Transform(M_CameratoWorld);
        return ray;
    }
};
struct PerspectiveCamera : CameraBase {            // :248-311
    float N, F, fov;
    PerspectiveCamera(float near_, float far_, float /*sw*/, float /*sh*/, float fov_, vec3 pos, vec3 look, vec3 right, vec3 up, vec2 res,
                      float lensR = 0, float focalD = 0)
        : CameraBase(2 * near_ * std::tan(radians(fov_) / 2.0f), 2 * near_ * std::tan(radians(fov_) / 2.0f) * (res.x / (float)res.y), pos, look,
                     right, up, res, lensR, focalD), N(near_), F(far_), fov(fov_) {
        calculuateMatrix();
    }
    void ChangeFOV(float f) { fov = f; calculuateMatrix(); }
    void calculuateMatrix() {                      // :303-310
        float invTanAng = 1.0f / std::tan(radians(fov) / 2.0f);
        mat4 persp = mat4::from_cols(vec4(1, 0, 0, 0), vec4(0, 1, 0, 0), vec4(0, 0, F / (F - N), 1), vec4(0, 0, -(F * N / (F - N)), 0));
        M_CameratoScreen = scale(persp, vec3(invTanAng, invTanAng, 1));
        M_RastertoCamera = mul(inverse(M_CameratoScreen), M_RastertoScreen);
    }
    Ray generateRay(vec2 pixel, Sampler* sampler) override {   // :273-297
        vec4 npw = mul(M_RastertoCamera, vec4(pixel.x, pixel.y, 0, 1));
        vec3 near_pos(npw.x / npw.w, npw.y / npw.w, npw.z / npw.w);
        Ray ray(vec3(0, 0, 0), normalize(near_pos));
        if (lensRadius > 0 && sampler) {
            vec2 lens_pos = vec2(lensRadius, lensRadius) * SampleUniformDiskConcentric(sampler->Get2D());
            float ft = focalDistance / ray.d.z;
            vec3 pfocus = ray.o + ray.d * ft;
            ray.o = vec3(lens_pos.x, lens_pos.y, 0);
            ray.d = normalize(pfocus - ray.o);
        }
        ray.Transform(M_CameratoWorld);
        return ray;
    }
};
struct PinholeCamera : CameraBase {                // :313-359
    float hole_radius; vec3 box_dimensions;
    PinholeCamera(float radius, vec3 box, vec3 pos, vec3 look, vec3 right, vec3 up, vec2 res)
        : CameraBase(box.x, box.y, pos, look, right, up, res), hole_radius(radius), box_dimensions(box) {}
    Ray generateRay(vec2 pixel, Sampler*) override {
        vec3 sensor_pos = xyz(mul(M_RastertoScreen, vec4(pixel.x, pixel.y, 0, 1)));
        vec3 pinhole(0.0f * hole_radius * std::cos(radians(0.0f)), 0.0f * hole_radius * std::sin(radians(0.0f)), box_dimensions.z);
        Ray ray(sensor_pos, normalize(pinhole - sensor_pos));
        ray.Transform(M_CameratoWorld);
        return ray;
    }
};
// A camera given directly by its two device-visible matrices (what the C ABI carries).
struct MatrixPerspectiveCamera : CameraBase {
    MatrixPerspectiveCamera(const mat4& r2c, const mat4& c2w, float lensR, float focalD)
        : CameraBase(1, 1, vec3(0, 0, 0), vec3(0, 0, 1), vec3(1, 0, 0), vec3(0, 1, 0), vec2(1, 1), lensR, focalD) {
        M_RastertoCamera = r2c;
        M_CameratoWorld = c2w;
    }
    Ray generateRay(vec2 pixel, Sampler* sampler) override {
        vec4 npw = mul(M_RastertoCamera, vec4(pixel.x, pixel.y, 0, 1));
        vec3 near_pos(npw.x / npw.w, npw.y / npw.w, npw.z / npw.w);
        Ray ray(vec3(0, 0, 0), normalize(near_pos));
        if (lensRadius > 0 && sampler) {
            vec2 lens_pos = vec2(lensRadius, lensRadius) * SampleUniformDiskConcentric(sampler->Get2D());
            float ft = focalDistance / ray.d.z;
            vec3 pfocus = ray.o + ray.d * ft;
            ray.o = vec3(lens_pos.x, lens_pos.y, 0);
            ray.d = normalize(pfocus - ray.o);
        }
        ray.Transform(M_CameratoWorld);
        return ray;
    }
};

// ---- Film.h:6-20 -------------------------------------------------------------------------------
struct pixel { vec3 rgbsum{0, 0, 0}; float weightsum = 0; };
struct Film {
    std::vector<pixel> pixels;
    ivec2 film_dim, image_res;
    PixelSensor* pixel_sensor = nullptr;
    Filter* filter = nullptr;
};

// ======================= Tier B scene description (oracle-defined) ===============================
enum MaterialType { MAT_LAMBERT = 0, MAT_DIELECTRIC = 1, MAT_CONDUCTOR = 2 };
struct Material {
    int type = MAT_LAMBERT;
    int refl = -1;          // reflectance spectrum id (Lambert); <0 = black (path ends)
    int eta = -1, k = -1;   // index of refraction / extinction spectrum ids
    int emit = -1;          // emission spectrum id; <0 = not emissive
    float emit_scale = 0;
    int two_sided = 0;
    bool eta_constant = true;
};
struct SurfaceHit {
    bool found = false;
    int kind = 0;           // 0 triangle, 1 analytic shape
    int mesh_id = -1, tri_id = -1, shape_id = -1;
    float t = 0, b0 = 0, b1 = 0, b2 = 0;
    vec3 p, ng_ff, ns_ff;   // geometric / shading normal, both flipped to face the incoming ray
    bool backside = false;  // the ray arrived from behind the outward normal
    int material = 0;
};
struct EmissiveTri { int mesh_id, tri_id; vec3 p0, p1, p2, n; float area; int material; };
// Lights.h:5-8 (comment-only in the reference): "points light: position, color, and r^2 falloff" / "sunlight: direction, color".
// kind 0 = point light at `v` with radiant intensity scale * spectrum (W/sr per nm): irradiance I cos / r^2;
// kind 1 = sun: `v` = unit direction TOWARDS the light, irradiance scale * spectrum on a surface facing it (no falloff: a direction has no distance).
struct DeltaLight { int kind = 0; vec3 v; int spectrum = -1; float scale = 1; };

struct IntegratorConfig {
    int mode = 0;           // 0 = Tier A reference Li (RayTracerTestApp.h:218-284), 1 = Tier B path
    int max_depth = 5;
    int rr_depth = 0;       // 0 = no Russian roulette
    float ray_eps = 1e-2f;
    float shadow_eps = 1e-3f;
    float albedo_rgb[3] = {0.5f, 0.5f, 0.5f};   // Tier A "colors" (RayTracerTestApp.h:207)
    // how the emissive triangles are sampled at a Lambert hit: 0 = one sample, light picked by the power CDF;
    // 1 = "1 sample from each light source" (Shading.h:4).  Point / sun lights are always sampled once each.
    int light_strategy = 0;
};

struct PathCounters { uint64_t paths = 0, closest_rays = 0, shadow_rays = 0, depth_sum = 0; };

struct Scene {
    Octtree_Model* oct = nullptr;
    std::vector<Shape*> shapes;
    std::vector<int> shape_material;
    std::vector<int> mesh_material;
    std::vector<Material> materials;
    std::vector<std::unique_ptr<Spectrum>> spectra;
    std::vector<EmissiveTri> lights;
    std::vector<DeltaLight> delta_lights;
    std::vector<float> light_cdf;
    float light_total = 0;

    void BuildLights();
    SurfaceHit Closest(const Ray& ray) const;
    bool Occluded(const Ray& ray, float tMax) const;
};

// ---- renderer ---------------------------------------------------------------------------------
struct Renderer {
    Scene* scene = nullptr;
    CameraBase* camera = nullptr;
    Film* film = nullptr;
    IntegratorConfig cfg;
    // Tier A constants built once (the reference rebuilds them per call, RayTracerTestApp.h:246-255)
    RGBIlluminantSpectrum lightA;
    RGBAlbedoSpectrum matA;

    void Prepare();
    SampledSpectrum LiReference(Ray ray, const SampledWavelengths& lambdas) const;
    SampledSpectrum LiPath(Ray ray, SampledWavelengths& lambdas, Sampler* sampler, PathCounters* pc) const;
    // evaluate_pixel, RayTracerTestApp.h:287-345.  Optional debug outputs for per-sample parity.
    struct SampleDebug { Ray ray; SampledWavelengths lambdas; SampledSpectrum L; vec3 rgb; float weight; };
    void evaluate_pixel(int pixel_id, int index, Sampler* sampl, SampleDebug* dbg = nullptr, PathCounters* pc = nullptr) const;
    // thread pool of RayTracerTestApp.h:349-409: static contiguous pixel ranges, one sampler per thread
    double RenderThreaded(const Sampler& proto, int spp_begin, int spp_end, int nthreads, int pixel_stride,
                          TraverseCounters* tc, PathCounters* pc) const;
};
void ResolveFilm(const Film& film, const mat3& RGBFromXYZ, unsigned char* out_rgb8, float* out_rgbf);   // :425-452

}  // namespace orc
