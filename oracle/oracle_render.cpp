// ORACLE (test infrastructure, never shipped, never on the product path).
// Renderer: Tier A = restatement of Applications/RayTracerTestApp.h:218-452;
// Tier B = oracle-defined path integrator (no reference implementation exists; "parity unpinned").
#include "oracle_render.h"

namespace orc {

bool g_faithful_overheads = false;
thread_local TraverseCounters* tl_counters = nullptr;

// ------------------------------------------------------------------ Tier B geometry helpers ------
static inline vec3 tri_pos(const MeshCache::Mesh& m, int tri, int k) { return m.positions[m.indices[3 * tri + k]]; }

void Scene::BuildLights() {
    lights.clear(); light_cdf.clear(); light_total = 0;
    if (!oct || materials.empty()) return;
    const auto& model = MeshCache::modelCache()[oct->model.model_name];
    for (size_t mi = 0; mi < model.meshes.size(); ++mi) {
        int mat = mesh_material[mi];
        if (materials[mat].emit < 0) continue;
        const auto& mesh = model.meshes[mi];
        for (size_t ti = 0; ti < mesh.indices.size() / 3; ++ti) {
            EmissiveTri e;
            e.mesh_id = (int)mi; e.tri_id = (int)ti; e.material = mat;
            e.p0 = tri_pos(mesh, (int)ti, 0); e.p1 = tri_pos(mesh, (int)ti, 1); e.p2 = tri_pos(mesh, (int)ti, 2);
            vec3 c = cross(e.p1 - e.p0, e.p2 - e.p0);
            float len = length(c);
            e.area = 0.5f * len;
            e.n = c * (1.0f / len);
            if (!(e.area > 0)) continue;
            lights.push_back(e);
            light_total += e.area * materials[mat].emit_scale;      // selection weight: power ~ area * scale
            light_cdf.push_back(light_total);
        }
    }
}

SurfaceHit Scene::Closest(const Ray& ray) const {
    SurfaceHit h;
    float tMax = std::numeric_limits<float>::max();
    if (oct) {
        auto rec = oct->TraverseClosest(ray);
        if (rec.found) {
            h.found = true; h.kind = 0; h.mesh_id = rec.info.mesh_id; h.tri_id = rec.info.tri_id;
            h.t = rec.isect.t; h.b0 = rec.isect.b0; h.b1 = rec.isect.b1; h.b2 = rec.isect.b2;
            tMax = h.t;
        }
    }
    int best_shape = -1;
    for (size_t s = 0; s < shapes.size(); ++s) {
        // analytic shapes after the mesh, in list order, strict '<' like the octree loop
        float t = -1;
        if (auto* sp = dynamic_cast<Sphere*>(shapes[s])) { auto is = sp->BasicIntersect(ray, tMax); if (is) t = is->t; }
        else if (auto* cy = dynamic_cast<Cylinder*>(shapes[s])) { auto is = cy->BasicIntersect(ray, tMax); if (is) t = is->t; }
        else if (auto* dk = dynamic_cast<Disk*>(shapes[s])) { auto is = dk->BasicIntersect(ray, tMax); if (is) t = is->t; }
        else if (auto* ts = dynamic_cast<TriangleSimple*>(shapes[s])) { auto is = ts->BasicIntersect(ray, tMax); if (is) t = is->t; }
        if (t >= 0 && t < tMax) { tMax = t; best_shape = (int)s; }
    }
    if (best_shape >= 0) {
        // Intersect() recomputes the same nearest root: a larger tMax can only un-reject, and this shape
        // already returned a hit under the smaller one.
        auto info = shapes[best_shape]->Intersect(ray, std::numeric_limits<float>::max());
        h = SurfaceHit();
        h.found = true; h.kind = 1; h.shape_id = best_shape; h.t = tMax;
        h.p = info->hitp;
        h.ns_ff = info->n; h.ng_ff = info->n;     // Shape::Intersect face-forwards n against the ray
        // was it flipped?  Redo the object-space test the shape used (Shapes.h:262-263 and siblings).
        vec3 n_obj(0, 0, 1), d_obj(0, 0, 1);
        if (auto* sp = dynamic_cast<Sphere*>(shapes[best_shape])) { auto is = sp->BasicIntersect(ray, std::numeric_limits<float>::max()); n_obj = normalize(vec3(2 * is->hitp.x, 2 * is->hitp.y, 2 * is->hitp.z)); d_obj = is->ray_d; }
        else if (auto* cy = dynamic_cast<Cylinder*>(shapes[best_shape])) { auto is = cy->BasicIntersect(ray, std::numeric_limits<float>::max()); n_obj = normalize(vec3(2 * is->hitp.x, 2 * is->hitp.y, 0)); d_obj = is->ray_d; }
        else if (auto* dk = dynamic_cast<Disk*>(shapes[best_shape])) { auto is = dk->BasicIntersect(ray, std::numeric_limits<float>::max()); n_obj = vec3(0, 0, 1); d_obj = is->ray_d; }
        else if (auto* ts = dynamic_cast<TriangleSimple*>(shapes[best_shape])) { auto is = ts->BasicIntersect(ray, std::numeric_limits<float>::max()); n_obj = normalize(cross(ts->p3 - ts->p1, ts->p2 - ts->p1)); d_obj = is->ray_d; }
        h.backside = dot(n_obj, d_obj) > 0;
        h.material = shape_material[best_shape];
        return h;
    }
    if (h.found) {
        const auto& mesh = MeshCache::modelCache()[oct->model.model_name].meshes[h.mesh_id];
        vec3 p0 = tri_pos(mesh, h.tri_id, 0), p1 = tri_pos(mesh, h.tri_id, 1), p2 = tri_pos(mesh, h.tri_id, 2);
        h.p = p0 * h.b0 + p1 * h.b1 + p2 * h.b2;
        vec3 ng = normalize(cross(p1 - p0, p2 - p0));
        vec3 ns = ng;
        if (!mesh.normals.empty()) {
            unsigned i0 = mesh.indices[3 * h.tri_id], i1 = mesh.indices[3 * h.tri_id + 1], i2 = mesh.indices[3 * h.tri_id + 2];
            ns = normalize(mesh.normals[i0] * h.b0 + mesh.normals[i1] * h.b1 + mesh.normals[i2] * h.b2);
        }
        h.backside = dot(ng, ray.d) > 0;
        h.ng_ff = h.backside ? -ng : ng;
        h.ns_ff = (dot(ns, ray.d) > 0) ? -ns : ns;      // same flip rule as Shapes.h:1074-1075
        h.material = mesh_material[h.mesh_id];
    }
    return h;
}

bool Scene::Occluded(const Ray& ray, float tMax) const {
    if (oct && oct->TraverseAny(ray, tMax)) return true;
    for (Shape* s : shapes) {
        if (auto* sp = dynamic_cast<Sphere*>(s)) { if (sp->BasicIntersect(ray, tMax)) return true; }
        else if (auto* cy = dynamic_cast<Cylinder*>(s)) { if (cy->BasicIntersect(ray, tMax)) return true; }
        else if (auto* dk = dynamic_cast<Disk*>(s)) { if (dk->BasicIntersect(ray, tMax)) return true; }
        else if (auto* ts = dynamic_cast<TriangleSimple*>(s)) { if (ts->BasicIntersect(ray, tMax)) return true; }
    }
    return false;
}

// ------------------------------------------------------------------ Tier A -----------------------
void Renderer::Prepare() {
    MakeRGBIlluminant(1, 1, 1, &lightA);                                       // RayTracerTestApp.h:246
    MakeRGBAlbedo(cfg.albedo_rgb[0], cfg.albedo_rgb[1], cfg.albedo_rgb[2], &matA);   // :254
    if (scene) scene->BuildLights();
}

SampledSpectrum Renderer::LiReference(Ray ray, const SampledWavelengths& lambdas) const {   // :218-284
    auto surf = scene->oct->Traverse(ray);
    if (surf.has_value()) {
        vec3 world_n = surf->n;
        SampledSpectrum radiance(0);
        SampledSpectrum light_spectral = lightA.Sample(lambdas);
        SampledSpectrum ambient_spectral = 0.3f * SpectraTables::get().illumF1->Sample(lambdas);
        SampledSpectrum mat_spectral = matA.Sample(lambdas);
        float light_1_cos = clampf(dot(world_n, vec3(0, 0, -1)), 0.0f, 1.0f);
        radiance += ambient_spectral;
        radiance += light_1_cos * (light_spectral * mat_spectral);
        return radiance;
    }
    return SampledSpectrum(0);
}

// ------------------------------------------------------------------ Tier B -----------------------
namespace {
struct cplx { float re, im; };
inline cplx cmul(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
inline cplx cadd(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
inline cplx csub(cplx a, cplx b) { return {a.re - b.re, a.im - b.im}; }
inline cplx cscale(float s, cplx a) { return {s * a.re, s * a.im}; }
inline cplx cdiv(cplx a, cplx b) {
    float scale = 1 / (b.re * b.re + b.im * b.im);
    return {scale * (a.re * b.re + a.im * b.im), scale * (a.im * b.re - a.re * b.im)};
}
inline float cnorm(cplx a) { return a.re * a.re + a.im * a.im; }
inline float cabs(cplx a) { return std::sqrt(cnorm(a)); }
inline cplx csqrt(cplx z) {    // pbrt-v4 util/math.h complex sqrt (no trig)
    float n = cabs(z), t1 = std::sqrt(.5f * (n + std::abs(z.re))), t2 = .5f * z.im / t1;
    if (n == 0) return {0, 0};
    if (z.re >= 0) return {t1, t2};
    return {std::abs(t2), std::copysign(t1, z.im)};
}
inline float FrDielectric(float cosTheta_i, float eta) {
    cosTheta_i = clampf(cosTheta_i, -1, 1);
    if (cosTheta_i < 0) { eta = 1 / eta; cosTheta_i = -cosTheta_i; }
    float sin2Theta_i = 1 - cosTheta_i * cosTheta_i;
    float sin2Theta_t = sin2Theta_i / (eta * eta);
    if (sin2Theta_t >= 1) return 1.f;
    float cosTheta_t = SafeSqrt(1 - sin2Theta_t);
    float r_parl = (eta * cosTheta_i - cosTheta_t) / (eta * cosTheta_i + cosTheta_t);
    float r_perp = (cosTheta_i - eta * cosTheta_t) / (cosTheta_i + eta * cosTheta_t);
    return (r_parl * r_parl + r_perp * r_perp) / 2;
}
inline float FrComplex(float cosTheta_i, float eta, float k) {
    cosTheta_i = clampf(cosTheta_i, 0, 1);
    cplx em{eta, k};
    float sin2Theta_i = 1 - cosTheta_i * cosTheta_i;
    cplx sin2Theta_t = cdiv(cplx{sin2Theta_i, 0}, cmul(em, em));
    cplx cosTheta_t = csqrt(csub(cplx{1, 0}, sin2Theta_t));
    cplx ec = cscale(cosTheta_i, em);
    cplx r_parl = cdiv(csub(ec, cosTheta_t), cadd(ec, cosTheta_t));
    cplx ect = cmul(em, cosTheta_t);
    cplx r_perp = cdiv(csub(cplx{cosTheta_i, 0}, ect), cadd(cplx{cosTheta_i, 0}, ect));
    return (cnorm(r_parl) + cnorm(r_perp)) / 2;
}
inline void CoordinateSystem(vec3 n, vec3* t, vec3* b) {      // pbrt-v4 (Duff et al.)
    float sign = std::copysign(1.0f, n.z);
    float a = -1 / (sign + n.z);
    float bb = n.x * n.y * a;
    *t = vec3(1 + sign * (n.x * n.x) * a, sign * bb, -sign * n.x);
    *b = vec3(bb, sign + (n.y * n.y) * a, -n.y);
}
inline vec3 OffsetOrigin(vec3 p, vec3 ng, vec3 w, float eps) {
    vec3 n = (dot(ng, w) < 0) ? -ng : ng;
    return p + n * eps;
}
}  // namespace

SampledSpectrum Renderer::LiPath(Ray ray, SampledWavelengths& lambdas, Sampler* sampler, PathCounters* pc) const {
    const Scene& sc = *scene;
    SampledSpectrum L(0), beta(1);
    bool specularBounce = true;
    int depth = 0;
    while (true) {
        if (pc) pc->closest_rays++;
        SurfaceHit h = sc.Closest(ray);
        if (!h.found) break;
        const Material& m = sc.materials[h.material];
        const bool backside = h.backside;
        const vec3 ng_ff = h.ng_ff, ns_ff = h.ns_ff;
        if (m.emit >= 0 && specularBounce && (m.two_sided || !backside))
            L += beta * (m.emit_scale * sc.spectra[m.emit]->Sample(lambdas));
        if (depth++ == cfg.max_depth) break;
        vec3 wo = -ray.d;
        vec3 wi;
        if (m.type == MAT_LAMBERT) {
            if (m.refl < 0) break;
            SampledSpectrum R = sc.spectra[m.refl]->Sample(lambdas);
            // next-event estimation over the emissive triangles: one sample of the light the power CDF picks (strategy 0), or
            // "1 sample from each light source" (Shading.h:4; strategy 1: every triangle with probability 1, in list order)
            auto sample_triangle = [&](const EmissiveTri& e, float pmf, vec2 up) {
                const Material& lm = sc.materials[e.material];
                float b0, b1;
                if (up.x < up.y) { b0 = up.x / 2; b1 = up.y - b0; } else { b1 = up.y / 2; b0 = up.x - b1; }
                float b2 = 1 - b0 - b1;
                vec3 pl = e.p0 * b0 + e.p1 * b1 + e.p2 * b2;
                vec3 so = OffsetOrigin(h.p, ng_ff, pl - h.p, cfg.ray_eps);
                vec3 dvec = pl - so;
                float dist2 = dot(dvec, dvec);
                float dist = std::sqrt(dist2);
                vec3 wl = dvec * (1.0f / dist);
                float cos_l = dot(e.n, -wl);
                if (lm.two_sided) cos_l = std::abs(cos_l);
                float cos_s = dot(ns_ff, wl);
                if (cos_l > 0 && cos_s > 0 && dot(ng_ff, wl) > 0) {
                    if (pc) pc->shadow_rays++;
                    if (!sc.Occluded(Ray(so, wl), dist * (1 - cfg.shadow_eps))) {
                        float pdf = pmf * dist2 / (e.area * cos_l);
                        SampledSpectrum Le = lm.emit_scale * sc.spectra[lm.emit]->Sample(lambdas);
                        L += beta * (R * InvPi) * Le * (cos_s / pdf);
                    }
                }
            };
            if (!sc.lights.empty() && cfg.light_strategy == 0) {
                float ul = sampler->Get1D();
                vec2 up = sampler->Get2D();
                float x = ul * sc.light_total;
                size_t lo = 0, hi = sc.light_cdf.size();
                while (lo < hi) { size_t mid = (lo + hi) / 2; if (sc.light_cdf[mid] > x) hi = mid; else lo = mid + 1; }
                size_t li = std::min(lo, sc.light_cdf.size() - 1);
                const EmissiveTri& e = sc.lights[li];
                float w_li = e.area * sc.materials[e.material].emit_scale;
                sample_triangle(e, w_li / sc.light_total, up);
            } else if (cfg.light_strategy == 1) {
                for (const EmissiveTri& e : sc.lights) sample_triangle(e, 1.0f, sampler->Get2D());
            }
            // point and sun lights (Lights.h:5-8): one deterministic sample each, in list order
            for (const DeltaLight& dl : sc.delta_lights) {
                vec3 so, wl;
                float tmax, atten;
                if (dl.kind == 0) {
                    so = OffsetOrigin(h.p, ng_ff, dl.v - h.p, cfg.ray_eps);
                    vec3 dvec = dl.v - so;
                    float dist2 = dot(dvec, dvec);
                    float dist = std::sqrt(dist2);
                    wl = dvec * (1.0f / dist);
                    tmax = dist * (1 - cfg.shadow_eps);
                    atten = 1.0f / dist2;                      // "r^2 falloff"
                } else {
                    wl = dl.v;
                    so = OffsetOrigin(h.p, ng_ff, wl, cfg.ray_eps);
                    tmax = std::numeric_limits<float>::max();
                    atten = 1.0f;
                }
                float cos_s = dot(ns_ff, wl);
                if (cos_s > 0 && dot(ng_ff, wl) > 0) {
                    if (pc) pc->shadow_rays++;
                    if (!sc.Occluded(Ray(so, wl), tmax)) {
                        SampledSpectrum I = dl.scale * sc.spectra[dl.spectrum]->Sample(lambdas);
                        L += beta * (R * InvPi) * I * (cos_s * atten);
                    }
                }
            }
            vec2 u = sampler->Get2D();
            vec3 wloc = SampleCosineHemisphere(u);
            if (wloc.z == 0) break;
            vec3 tx, ty;
            CoordinateSystem(ns_ff, &tx, &ty);
            wi = tx * wloc.x + ty * wloc.y + ns_ff * wloc.z;
            if (!(dot(wi, ng_ff) > 0)) break;
            beta *= R;
            specularBounce = false;
        } else if (m.type == MAT_DIELECTRIC) {
            float eta = sc.spectra[m.eta]->Query(lambdas.lambda[0]);
            if (!m.eta_constant) lambdas.TerminateSecondary();
            vec3 n = ns_ff;
            bool entering = !backside;
            float etap = entering ? eta : 1 / eta;
            float cos_i = dot(wo, n);
            float Rf = FrDielectric(cos_i, etap);
            float uc = sampler->Get1D();
            if (uc < Rf) {
                wi = -wo + n * (2 * dot(wo, n));
            } else {
                float sin2_i = std::max(0.f, 1 - cos_i * cos_i);
                float sin2_t = sin2_i / (etap * etap);
                if (sin2_t >= 1) break;
                float cos_t = SafeSqrt(1 - sin2_t);
                wi = -wo / etap + n * (cos_i / etap - cos_t);
                beta *= 1 / (etap * etap);
            }
            specularBounce = true;
        } else {   // MAT_CONDUCTOR
            vec3 n = ns_ff;
            float cos_i = dot(wo, n);
            SampledSpectrum e = sc.spectra[m.eta]->Sample(lambdas), k = sc.spectra[m.k]->Sample(lambdas);
            SampledSpectrum F;
            for (int i = 0; i < NSpectrumSamples; ++i) F[i] = FrComplex(cos_i, e[i], k[i]);
            wi = -wo + n * (2 * cos_i);
            beta *= F;
            specularBounce = true;
        }
        if (cfg.rr_depth > 0 && depth >= cfg.rr_depth) {
            float mx = beta.MaxComponentValue();
            if (mx < 1) {
                float q = std::max(0.f, 1 - mx);
                if (sampler->Get1D() < q) break;
                beta *= 1 / (1 - q);
            }
        }
        wi = normalize(wi);
        ray = Ray(OffsetOrigin(h.p, ng_ff, wi, cfg.ray_eps), wi);
    }
    if (pc) { pc->paths++; pc->depth_sum += depth; }
    return L;
}

// ------------------------------------------------------------------ evaluate_pixel ----------------
void Renderer::evaluate_pixel(int pixel_id, int index, Sampler* sampl, SampleDebug* dbg, PathCounters* pc) const {   // :287-345
    int x_pix = pixel_id % film->image_res.x;
    int y_pix = (int)(film->image_res.y - std::floor(pixel_id / (float)film->image_res.x));
    ivec2 pix(x_pix, y_pix);
    sampl->StartPixelSample(pix, index, 0);
    SampledWavelengths lambdas = SampledWavelengths::SampleVisible(sampl->Get1D());
    vec2 uniform_pixel_offset = sampl->GetPixel2D();
    FilterSample fs = film->filter->Sample(uniform_pixel_offset);
    vec2 pixel_sampled_pos = vec2((float)pix.x, (float)pix.y) + vec2(.5f, .5f) + fs.p;
    Ray ray = camera->generateRay(pixel_sampled_pos, sampl);
    if (dbg) { dbg->ray = ray; dbg->lambdas = lambdas; }
    SampledSpectrum L = (cfg.mode == 0) ? LiReference(ray, lambdas) : LiPath(ray, lambdas, sampl, pc);
    if (cfg.mode == 0 && pc) { pc->paths++; pc->closest_rays++; }
    vec3 cam = film->pixel_sensor->ToSensorRGB(L, lambdas);
    cam.x = clampf(cam.x, 0.0f, 1.0f);
    cam.y = clampf(cam.y, 0.0f, 1.0f);
    cam.z = clampf(cam.z, 0.0f, 1.0f);
    film->pixels[pixel_id].rgbsum += fs.weight * cam;
    film->pixels[pixel_id].weightsum += fs.weight;
    if (dbg) { dbg->L = L; dbg->rgb = cam; dbg->weight = fs.weight; dbg->lambdas = lambdas; }
}

double Renderer::RenderThreaded(const Sampler& proto, int spp_begin, int spp_end, int nthreads, int pixel_stride,
                                TraverseCounters* tc, PathCounters* pc) const {   // :349-409
    int npix = (int)film->pixels.size();
    int thread_count = std::max(nthreads, 1);
    int per = npix / thread_count;
    std::vector<TraverseCounters> tcs(thread_count);
    std::vector<PathCounters> pcs(thread_count);
    std::vector<std::thread> pool;
    auto t1 = std::chrono::high_resolution_clock::now();
    int begin = 0;
    for (int i = 0; i < thread_count; ++i) {
        int end = begin + per;
        if (i == thread_count - 1) end = npix;
        pool.emplace_back([&, i, begin, end] {
            std::unique_ptr<Sampler> s = proto.Clone();
            tl_counters = tc ? &tcs[i] : nullptr;
            for (int idx = spp_begin; idx < spp_end; ++idx)
                for (int j = begin; j < end; ++j) {
                    if (pixel_stride > 1 && (j % pixel_stride) != 0) continue;
                    evaluate_pixel(j, idx, s.get(), nullptr, pc ? &pcs[i] : nullptr);
                }
            tl_counters = nullptr;
        });
        begin = end;
    }
    for (auto& t : pool) t.join();
    auto t2 = std::chrono::high_resolution_clock::now();
    if (tc) for (auto& c : tcs) tc->add(c);
    if (pc) for (auto& c : pcs) { pc->paths += c.paths; pc->closest_rays += c.closest_rays; pc->shadow_rays += c.shadow_rays; pc->depth_sum += c.depth_sum; }
    return std::chrono::duration<double>(t2 - t1).count();
}

void ResolveFilm(const Film& film, const mat3& RGBFromXYZ, unsigned char* out8, float* outf) {   // :425-452
    for (size_t i = 0; i < film.pixels.size(); ++i) {
        vec3 sensor_rgb = film.pixels[i].rgbsum / film.pixels[i].weightsum;
        vec3 xyz_val = mul(film.pixel_sensor->XYZFromSensorRGB, sensor_rgb);
        vec3 o = mul(RGBFromXYZ, xyz_val);
        o.x = clampf(o.x, 0.0f, 1.0f); o.y = clampf(o.y, 0.0f, 1.0f); o.z = clampf(o.z, 0.0f, 1.0f);
        if (outf) { outf[3 * i] = o.x; outf[3 * i + 1] = o.y; outf[3 * i + 2] = o.z; }
        if (out8) {
            float scale = 255.0f;
            out8[3 * i] = (unsigned char)(scale * o.x);
            out8[3 * i + 1] = (unsigned char)(scale * o.y);
            out8[3 * i + 2] = (unsigned char)(scale * o.z);
        }
    }
}

}  // namespace orc
