// crt_path.cuh -- wavefront path integrator stages (Tier B): the integrator the reference only names
// (RayTracer/Integrator.h:4-12 "PathIntegrator"), shaded the way Shading.h:1-20 and Lights.h:1-9 sketch:
// Lambert r/pi with one light sample per hit (next-event estimation over emissive triangles), perfect
// reflect / refract weighted by Fresnel, conductors, optional Russian roulette.  There is no reference code
// for these stages; their definition is oracle/oracle_render.cpp (Renderer::LiPath, Scene::Closest,
// Scene::Occluded) and every expression below follows that file's operation order.
//
// One wave = one sample index over the owned pixels.  Per bounce:
//   k_trace<closest>  over the active queue            (crt_trace.cuh / crt_kernels.cuh)
//   k_path_shade      surface record, emission, light sample -> shadow queue, BSDF sample -> next active queue
//   k_trace<any>      over the shadow queue
//   k_shadow_resolve  adds the unoccluded light contributions to the path radiance
// and k_path_splat at the end of the wave (ToSensorRGB + clamp + film accumulate, RayTracerTestApp.h:326-337).
// Queues are compacted with one atomic per warp (__ballot_sync + __popc).
#pragma once
#include "crt_kernels.cuh"

namespace crt {

#define CRT_FLAG_SPECULAR 1u
#define CRT_INV_PI 0.31830988618379067154f

struct SurfaceHitDev {
    int found, kind, id0, id1, material, backside;
    float t;
    f3 p, ng_ff, ns_ff;
};

// Scene::Closest, triangle branch (oracle_render.cpp:77-91)
CRT_D void surface_from_triangle(const DeviceScene& S, int ref, float4 tb, f3 rd, SurfaceHitDev& h) {
    float4 v0 = S.tris[3 * (size_t)ref], v1 = S.tris[3 * (size_t)ref + 1], v2 = S.tris[3 * (size_t)ref + 2];
    f3 p0 = mk3(v0.x, v0.y, v0.z), p1 = mk3(v1.x, v1.y, v1.z), p2 = mk3(v2.x, v2.y, v2.z);
    h.found = 1; h.kind = 0;
    h.material = __float_as_int(v0.w); h.id0 = __float_as_int(v1.w); h.id1 = __float_as_int(v2.w);
    h.t = tb.x;
    h.p = (p0 * tb.y + p1 * tb.z) + p2 * tb.w;
    f3 ng = normalize3(cross3(p1 - p0, p2 - p0));
    f3 ns = ng;
    if (S.tri_nrm) {
        float4 n0 = S.tri_nrm[3 * (size_t)ref], n1 = S.tri_nrm[3 * (size_t)ref + 1], n2 = S.tri_nrm[3 * (size_t)ref + 2];
        ns = normalize3((mk3(n0.x, n0.y, n0.z) * tb.y + mk3(n1.x, n1.y, n1.z) * tb.z) + mk3(n2.x, n2.y, n2.z) * tb.w);
    }
    h.backside = dot3(ng, rd) > 0;
    h.ng_ff = h.backside ? -ng : ng;
    h.ns_ff = (dot3(ns, rd) > 0) ? -ns : ns;
}

// Scene::Closest, analytic shapes after the mesh, in list order, strict '<' (oracle_render.cpp:49-76).
// The list-order loop returns the shape with the lexicographically smallest (t, index) among those nearer than the mesh hit: a shape's
// BasicIntersect(ray, tMax) is its first valid root truncated at tMax, so the order of the tests matters only for exact ties.  The shapes
// are therefore visited through a threaded BVH over their padded boxes (skip pointers, no stack) with that tie rule spelled out.
CRT_D void closest_over_shapes(const DeviceScene& S, f3 ro, f3 rd, SurfaceHitDev& h) {
    float tMax = h.found ? h.t : FLT_MAX;
    int best = -1;
    f3 best_p = mk3(0, 0, 0), best_d = best_p;          // object-space hit point and direction of the best shape so far
    RayConst rb;
    rb.o = ro; rb.inv_d = mk3(1 / rd.x, 1 / rd.y, 1 / rd.z);
    int i = 0;
    while (true) {
        // walk the boxes up to the next shape whose box the ray enters within tMax; the lanes of a warp meet again at the shape test
        int s = -1;
        while (i < S.n_shape_nodes) {
            const float4 lo = __ldg(&S.shape_bvh[2 * i]), hi = __ldg(&S.shape_bvh[2 * i + 1]);
            float m;
            if (!slab_unbounded(rb, lo, hi, m) || m > tMax) { i = __float_as_int(lo.w); continue; }      // cannot be hit within tMax: skip the subtree
            ++i;
            s = __float_as_int(hi.w);
            if (s >= 0) break;
        }
        if (s < 0) break;
        const bool may_tie = best >= 0 && s < best;          // an equal t of a lower-numbered shape wins in list order
        ShapeIsect is;
        if (shape_basic_lean(S.shapes[s], ro, rd, may_tie ? nextafterf(tMax, INFINITY) : tMax, is)) {
            if (is.t >= 0 && (is.t < tMax || (may_tie && is.t == tMax))) { tMax = is.t; best = s; best_p = is.hitp; best_d = is.ray_d; }
        }
    }
    if (best < 0) return;
    // (BasicIntersect returns a shape's first valid root whatever tMax it is given, as long as that root is within tMax: the record kept
    // at the accepting test is the one Shape::Intersect(ray, FLT_MAX) would form)
    int flipped;
    shape_surface_lean(S.shapes[best], best_p, best_d, h.p, h.ns_ff, flipped);
    h.found = 1; h.kind = 1; h.id0 = best; h.id1 = -1; h.t = tMax;
    h.ng_ff = h.ns_ff;
    h.backside = flipped;
    h.material = S.shapes[best].material;
}
CRT_D bool occluded_by_shapes(const DeviceScene& S, f3 ro, f3 rd, float tMax) {
    RayConst rb;
    rb.o = ro; rb.inv_d = mk3(1 / rd.x, 1 / rd.y, 1 / rd.z);
    int i = 0;
    while (true) {
        int s = -1;
        while (i < S.n_shape_nodes) {
            const float4 lo = __ldg(&S.shape_bvh[2 * i]), hi = __ldg(&S.shape_bvh[2 * i + 1]);
            float m;
            if (!slab_unbounded(rb, lo, hi, m) || m > tMax) { i = __float_as_int(lo.w); continue; }
            ++i;
            s = __float_as_int(hi.w);
            if (s >= 0) break;
        }
        if (s < 0) return false;
        ShapeIsect is;
        if (shape_basic_lean(S.shapes[s], ro, rd, tMax, is)) return true;
    }
}

// ---- BSDF helpers (oracle_render.cpp:130-182; pbrt-v4 formulas) -----------------------------------------
struct cplx { float re, im; };
CRT_D cplx c_mul(cplx a, cplx b) { cplx r; r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; return r; }
CRT_D cplx c_add(cplx a, cplx b) { cplx r; r.re = a.re + b.re; r.im = a.im + b.im; return r; }
CRT_D cplx c_sub(cplx a, cplx b) { cplx r; r.re = a.re - b.re; r.im = a.im - b.im; return r; }
CRT_D cplx c_scale(float s, cplx a) { cplx r; r.re = s * a.re; r.im = s * a.im; return r; }
CRT_D cplx c_div(cplx a, cplx b) {
    float scale = 1 / (b.re * b.re + b.im * b.im);
    cplx r; r.re = scale * (a.re * b.re + a.im * b.im); r.im = scale * (a.im * b.re - a.re * b.im);
    return r;
}
CRT_D float c_norm(cplx a) { return a.re * a.re + a.im * a.im; }
CRT_D cplx c_sqrt(cplx z) {
    float n = sqrtf(c_norm(z)), t1 = sqrtf(.5f * (n + fabsf(z.re))), t2 = .5f * z.im / t1;
    cplx r;
    if (n == 0) { r.re = 0; r.im = 0; return r; }
    if (z.re >= 0) { r.re = t1; r.im = t2; return r; }
    r.re = fabsf(t2); r.im = copysignf(t1, z.im);
    return r;
}
CRT_D float fr_dielectric(float cosTheta_i, float eta) {
    cosTheta_i = gclamp(cosTheta_i, -1.f, 1.f);
    if (cosTheta_i < 0) { eta = 1 / eta; cosTheta_i = -cosTheta_i; }
    float sin2Theta_i = 1 - cosTheta_i * cosTheta_i;
    float sin2Theta_t = sin2Theta_i / (eta * eta);
    if (sin2Theta_t >= 1) return 1.f;
    float cosTheta_t = safe_sqrt(1 - sin2Theta_t);
    float r_parl = (eta * cosTheta_i - cosTheta_t) / (eta * cosTheta_i + cosTheta_t);
    float r_perp = (cosTheta_i - eta * cosTheta_t) / (cosTheta_i + eta * cosTheta_t);
    return (r_parl * r_parl + r_perp * r_perp) / 2;
}
// (not inlined: eight copies per conductor hit, see spectrum_query_nl)
__device__ __noinline__ float fr_complex(float cosTheta_i, float eta, float k) {
    cosTheta_i = gclamp(cosTheta_i, 0.f, 1.f);
    cplx em; em.re = eta; em.im = k;
    float sin2Theta_i = 1 - cosTheta_i * cosTheta_i;
    cplx s2; s2.re = sin2Theta_i; s2.im = 0;
    cplx one; one.re = 1; one.im = 0;
    cplx ci; ci.re = cosTheta_i; ci.im = 0;
    cplx sin2Theta_t = c_div(s2, c_mul(em, em));
    cplx cosTheta_t = c_sqrt(c_sub(one, sin2Theta_t));
    cplx ec = c_scale(cosTheta_i, em);
    cplx r_parl = c_div(c_sub(ec, cosTheta_t), c_add(ec, cosTheta_t));
    cplx ect = c_mul(em, cosTheta_t);
    cplx r_perp = c_div(c_sub(ci, ect), c_add(ci, ect));
    return (c_norm(r_parl) + c_norm(r_perp)) / 2;
}
CRT_D void coordinate_system(f3 n, f3& t, f3& b) {          // pbrt-v4 (Duff et al.)
    float sign = copysignf(1.0f, n.z);
    float a = -1 / (sign + n.z);
    float bb = n.x * n.y * a;
    t = mk3(1 + sign * (n.x * n.x) * a, sign * bb, -sign * n.x);
    b = mk3(bb, sign + (n.y * n.y) * a, -n.y);
}
CRT_D f3 offset_origin(f3 p, f3 ng, f3 w, float eps) {
    f3 n = (dot3(ng, w) < 0) ? -ng : ng;
    return p + n * eps;
}
// SampleCosineHemisphere (RayTracer/Sampling.h:449-454)
CRT_D f3 sample_cosine_hemisphere(f2 u) {
    f2 d = sample_disk_concentric(u);
    float z = safe_sqrt(1 - d.x * d.x - d.y * d.y);
    return mk3(d.x, d.y, z);
}

// ---- queues -------------------------------------------------------------------------------------------
struct PathQueues {
    const int* active;        // path ids of this bounce (nullptr = identity over n)
    const int* n_active;      // device count (nullptr = n)
    int n;
    int* next_active; int* n_next;
    // shadow queue of this bounce: slot -> ray, contribution, owning path
    float4* sh_o; float4* sh_d; float4* sh_k; float4* sh_s; float4* sh_contrib; int* sh_path; int* n_shadow;
    unsigned long long* ray_counters;   // [0] closest rays, [1] shadow rays, [2] depth sum
    int count_active;                   // k_shadow_resolve also adds this bounce's closest-ray count (once per bounce)
};

// one atomic per warp: returns this lane's slot (valid only where pred), __ballot_sync + __popc compaction
CRT_D int warp_enqueue(bool pred, int* counter) {
    unsigned active = __activemask();
    unsigned m = __ballot_sync(active, pred);
    if (!m) return -1;
    int lane = threadIdx.x & 31;
    int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(active, base, leader);
    return base + __popc(m & ((1u << lane) - 1));
}

struct PathDebugOut { int* kind; int* id0; int* id1; float* t; float* p3; float* ns3; float* ng3; int* backside; };

// Scene::Closest for path slot i from the traversal's hit record + the analytic shapes (oracle_render.cpp:38-93)
// the mesh part: the record of the traversal launch, or (root_leaf) the traversal itself
CRT_D void mesh_surface_hit(const DeviceScene& S, const PathBuffers& pb, int i, f3 ro, f3 rd, float tMax, SurfaceHitDev& h) {
    h.found = 0; h.kind = -1; h.id0 = -1; h.id1 = -1; h.material = 0; h.backside = 0; h.t = 0;
    h.p = mk3(0, 0, 0); h.ng_ff = h.p; h.ns_ff = h.p;
    int ref = -1;
    float4 tb = make_float4(0, 0, 0, 0);
    if (S.root_leaf) trace_root_leaf<false>(S, ro, rd, tMax, ref, tb);
    else if (S.has_model) { ref = pb.hit_ref[i]; if (ref >= 0) tb = pb.hit_tb[i]; }
    if (ref >= 0) surface_from_triangle(S, ref, tb, rd, h);
}
CRT_D void path_surface_hit(const DeviceScene& S, const PathBuffers& pb, int i, f3 ro, f3 rd, float tMax, SurfaceHitDev& h) {
    mesh_surface_hit(S, pb, i, ro, rd, tMax, h);
    if (S.n_shapes > 0) closest_over_shapes(S, ro, rd, h);
}

// one light sample at a Lambert hit -> shadow ray + the contribution it carries if unoccluded (oracle_render.cpp, sample_triangle)
struct LightSample { bool valid; float4 o, d; Spec8 contrib; };
CRT_D void sample_emissive_triangle(const DeviceScene& S, const RenderConst& rc, const SurfaceHitDev& h, const Spec8& lambda, const Spec8& beta, const Spec8& R,
                                    const DevLight& e, float pmf, f2 up, LightSample& ls) {
    const DevMaterial lm = S.materials[e.material];
    float b0, b1;
    if (up.x < up.y) { b0 = up.x / 2; b1 = up.y - b0; } else { b1 = up.y / 2; b0 = up.x - b1; }
    float b2 = 1 - b0 - b1;
    f3 e0 = mk3(e.p0[0], e.p0[1], e.p0[2]), e1 = mk3(e.p1[0], e.p1[1], e.p1[2]), e2 = mk3(e.p2[0], e.p2[1], e.p2[2]);
    f3 en = mk3(e.n[0], e.n[1], e.n[2]);
    f3 pl = (e0 * b0 + e1 * b1) + e2 * b2;
    f3 so = offset_origin(h.p, h.ng_ff, pl - h.p, rc.ray_eps);
    f3 dvec = pl - so;
    float dist2 = dot3(dvec, dvec);
    float dist = sqrtf(dist2);
    f3 wl = dvec * (1.0f / dist);
    float cos_l = dot3(en, -wl);
    if (lm.two_sided) cos_l = fabsf(cos_l);
    float cos_s = dot3(h.ns_ff, wl);
    ls.valid = false;
    if (cos_l > 0 && cos_s > 0 && dot3(h.ng_ff, wl) > 0) {
        float pdf = pmf * dist2 / (e.area * cos_l);
        float g = cos_s / pdf;
#pragma unroll
        for (int k = 0; k < CRT_NLAMBDA; ++k) {
            float Le = spectrum_query(S, lm.emit, lambda.v[k]) * lm.emit_scale;
            ls.contrib.v[k] = ((beta.v[k] * (R.v[k] * CRT_INV_PI)) * Le) * g;
        }
        ls.valid = true;
        ls.o = make_float4(so.x, so.y, so.z, dist * (1 - rc.shadow_eps));
        ls.d = make_float4(wl.x, wl.y, wl.z, 0);
    }
}

// Renderer::LiPath loop body for one bounce (oracle_render.cpp:189-285) of path slot i at the surface record h: emission, light sample,
// BSDF sample, Russian roulette; advances the path state in place and returns what goes into the queues.  MAT is the material type when
// the caller knows it at compile time (staged shading, one kernel per type) or -1 (read from the material).
struct ShadeOut { bool continues, has_shadow; float4 sh_o, sh_d; Spec8 contrib; };
template <int MAT>
CRT_D void shade_bounce(const DeviceScene& S, const RenderConst& rc, const PathBuffers& pb, int i, f3 rd, const SurfaceHitDev& h, ShadeOut& out) {
    unsigned flags = (unsigned)pb.flags[i];
    int depth = (int)(flags >> 8);
    bool specular = flags & CRT_FLAG_SPECULAR;
    const DevMaterial m = S.materials[h.material];
    const int mtype = MAT < 0 ? m.type : MAT;
    Spec8 lambda, pdfw, beta, L;
    load8(pb.lambda, i, lambda); load8(pb.beta, i, beta);
    // L and the wavelength pdfs are only touched by emissive hits / dispersive dielectrics: load them there
    bool L_dirty = false, pdf_dirty = false;
    if (m.emit >= 0 && specular && (m.two_sided || !h.backside)) {
        load8(pb.L, i, L);
#pragma unroll
        for (int k = 0; k < CRT_NLAMBDA; ++k) L.v[k] += beta.v[k] * (spectrum_query(S, m.emit, lambda.v[k]) * m.emit_scale);
        L_dirty = true;
    }
    bool go = !(depth++ == rc.max_depth);
    f3 wo = -rd, wi = mk3(0, 0, 1);
    SamplerState ss;
    if (go) ss = pb.sampler[i];
    if (go && mtype == MAT_LAMBERT) {
        if (m.refl < 0) go = false;
        Spec8 R;
        if (go) {
            spectrum_sample(S, m.refl, lambda, R);
            if (S.n_lights > 0 && rc.light_strategy == 0) {            // next-event estimation: one sample of the light the power CDF picks
                float ul = sampler_get1d(rc.sampler, ss);
                f2 up = sampler_get2d(rc.sampler, ss);
                float x = ul * S.light_total;
                int lo = 0, hi = S.n_lights;
                while (lo < hi) { int mid = (lo + hi) / 2; if (__ldg(&S.light_cdf[mid]) > x) hi = mid; else lo = mid + 1; }
                int li = min(lo, S.n_lights - 1);
                const DevLight e = S.lights[li];
                float w_li = e.area * S.materials[e.material].emit_scale;
                LightSample ls;
                sample_emissive_triangle(S, rc, h, lambda, beta, R, e, w_li / S.light_total, up, ls);
                if (ls.valid) { out.has_shadow = true; out.sh_o = ls.o; out.sh_d = ls.d; out.contrib = ls.contrib; }
            }
            f2 u = sampler_get2d(rc.sampler, ss);
            f3 wloc = sample_cosine_hemisphere(u);
            if (wloc.z == 0) go = false;
            if (go) {
                f3 tx, ty;
                coordinate_system(h.ns_ff, tx, ty);
                wi = (tx * wloc.x + ty * wloc.y) + h.ns_ff * wloc.z;
                if (!(dot3(wi, h.ng_ff) > 0)) go = false;
            }
            if (go) {
#pragma unroll
                for (int k = 0; k < CRT_NLAMBDA; ++k) beta.v[k] *= R.v[k];
                specular = false;
            }
        }
    } else if (go && mtype == MAT_DIELECTRIC) {
        float eta = spectrum_query(S, m.eta, lambda.v[0]);
        if (!m.eta_constant) {                   // SampledWavelengths::TerminateSecondary, spectrum.h:302-310
            load8(pb.pdf, i, pdfw);
            bool terminated = true;
#pragma unroll
            for (int k = 1; k < CRT_NLAMBDA; ++k) if (pdfw.v[k] != 0) terminated = false;
            if (!terminated) {
#pragma unroll
                for (int k = 1; k < CRT_NLAMBDA; ++k) pdfw.v[k] = 0;
                pdfw.v[0] /= CRT_NLAMBDA;
                pdf_dirty = true;
            }
        }
        f3 nn = h.ns_ff;
        bool entering = !h.backside;
        float etap = entering ? eta : 1 / eta;
        float cos_i = dot3(wo, nn);
        float Rf = fr_dielectric(cos_i, etap);
        float uc = sampler_get1d(rc.sampler, ss);
        if (uc < Rf) {
            wi = -wo + nn * (2 * dot3(wo, nn));
        } else {
            float sin2_i = max_std(0.f, 1 - cos_i * cos_i);
            float sin2_t = sin2_i / (etap * etap);
            if (sin2_t >= 1) go = false;
            else {
                float cos_t = safe_sqrt(1 - sin2_t);
                wi = -wo / etap + nn * (cos_i / etap - cos_t);
                float sc = 1 / (etap * etap);
#pragma unroll
                for (int k = 0; k < CRT_NLAMBDA; ++k) beta.v[k] *= sc;
            }
        }
        specular = true;
    } else if (go) {   // MAT_CONDUCTOR
        f3 nn = h.ns_ff;
        float cos_i = dot3(wo, nn);
#pragma unroll
        for (int k = 0; k < CRT_NLAMBDA; ++k) {
            float ev = spectrum_query(S, m.eta, lambda.v[k]), kv = spectrum_query(S, m.k, lambda.v[k]);
            beta.v[k] *= fr_complex(cos_i, ev, kv);
        }
        wi = -wo + nn * (2 * cos_i);
        specular = true;
    }
    if (go && rc.rr_depth > 0 && depth >= rc.rr_depth) {
        float mx = beta.v[0];
#pragma unroll
        for (int k = 1; k < CRT_NLAMBDA; ++k) mx = max_std(mx, beta.v[k]);
        if (mx < 1) {
            float q = max_std(0.f, 1 - mx);
            if (sampler_get1d(rc.sampler, ss) < q) go = false;
            else {
                float sc = 1 / (1 - q);
#pragma unroll
                for (int k = 0; k < CRT_NLAMBDA; ++k) beta.v[k] *= sc;
            }
        }
    }
    if (go) {
        wi = normalize3(wi);
        const f3 new_o = offset_origin(h.p, h.ng_ff, wi, rc.ray_eps);
        out.continues = true;
        store8(pb.beta, i, beta);
        pb.sampler[i] = ss;
        store_ray(pb.ray_o, pb.ray_d, pb.ray_k, pb.ray_s, i, new_o, wi, FLT_MAX);
    } else if (out.has_shadow) {
        // the path ends here but its light sample was drawn before the terminating test: the oracle has
        // already added it (oracle_render.cpp:229-234 precede :238,:242,:279)
    }
    if (L_dirty) store8(pb.L, i, L);
    if (pdf_dirty) store8(pb.pdf, i, pdfw);
    pb.flags[i] = (int)(((unsigned)depth << 8) | (specular ? CRT_FLAG_SPECULAR : 0u));
}

// queue compaction shared by the shading kernels (all lanes of the warp call it)
CRT_D void enqueue_bounce(const PathQueues& Q, int i, const ShadeOut& out, bool with_shadow) {
    if (with_shadow) {
        const int s_slot = warp_enqueue(out.has_shadow, Q.n_shadow);
        if (out.has_shadow) {
            store_ray(Q.sh_o, Q.sh_d, Q.sh_k, Q.sh_s, s_slot, mk3(out.sh_o.x, out.sh_o.y, out.sh_o.z), mk3(out.sh_d.x, out.sh_d.y, out.sh_d.z), out.sh_o.w);
            Q.sh_path[s_slot] = i;
            store8(Q.sh_contrib, s_slot, out.contrib);
        }
    }
    const int a_slot = warp_enqueue(out.continues, Q.n_next);
    if (out.continues) Q.next_active[a_slot] = i;
}

// ---- fused shading: surface record + bounce in one kernel, one thread per queue slot (scenes without analytic shapes) --------------
#ifndef CRT_SHADE_MINBLOCKS
#define CRT_SHADE_MINBLOCKS 7          // 73 registers (measured on C3, 5 / 6 / 7 / 8 CTAs per SM: 381 / 413 / 434 / 431 Mpaths/s)
#endif
__global__ void __launch_bounds__(128, CRT_SHADE_MINBLOCKS) k_path_shade(DeviceScene S, RenderConst rc, PathBuffers pb, PathQueues Q, PathDebugOut dbg) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = Q.n_active ? *Q.n_active : Q.n;
    const bool live = slot < n;
    int i = 0;
    ShadeOut out;
    out.continues = false; out.has_shadow = false; out.sh_o = make_float4(0, 0, 0, 0); out.sh_d = out.sh_o;
    if (live) {
        i = Q.active ? Q.active[slot] : slot;
        float4 o4 = pb.ray_o[i], d4 = pb.ray_d[i];
        f3 ro = mk3(o4.x, o4.y, o4.z), rd = mk3(d4.x, d4.y, d4.z);
        SurfaceHitDev h;
        path_surface_hit(S, pb, i, ro, rd, o4.w, h);
        if (dbg.kind) {
            dbg.kind[i] = h.found ? h.kind : -1; dbg.id0[i] = h.id0; dbg.id1[i] = h.id1; dbg.t[i] = h.t; dbg.backside[i] = h.backside;
            dbg.p3[3 * i] = h.p.x; dbg.p3[3 * i + 1] = h.p.y; dbg.p3[3 * i + 2] = h.p.z;
            dbg.ns3[3 * i] = h.ns_ff.x; dbg.ns3[3 * i + 1] = h.ns_ff.y; dbg.ns3[3 * i + 2] = h.ns_ff.z;
            dbg.ng3[3 * i] = h.ng_ff.x; dbg.ng3[3 * i + 1] = h.ng_ff.y; dbg.ng3[3 * i + 2] = h.ng_ff.z;
        }
        if (h.found && !dbg.kind) shade_bounce<-1>(S, rc, pb, i, rd, h, out);
    }
    enqueue_bounce(Q, i, out, true);
}

// ---- staged shading (scenes with analytic shapes) ------------------------------------------------------------------------------------
// The fused kernel above, with the shape hierarchy, the four shapes' intersection and surface code and three materials inlined into one
// body, is 6.8 k instructions (110 KB) and its warps diverge over all of it: ncu attributes 48 % of its stall samples to instruction
// fetch ("no instruction") and counts 9-20 active threads per instruction.  Staged, a bounce is
//   k_path_hit            surface record of every active path (triangle record or shape hierarchy) -> HitRecords, path id -> the queue of
//                         its material type
//   k_path_shade_mat<T>   one launch per material type over that type's queue: no divergence over materials, a code body that fits the
//                         instruction cache
// with persistent grids (a warp strides over the queue), so a nearly empty deep bounce costs a few microseconds instead of a full-size
// grid of CTAs that exit at once.  Per path the operations are those of the fused kernel in the same order: films are bit-identical
// (tests/test_gpu_path.py::test_staged_shading_renders_the_identical_film).
struct HitRecords { float4* a; float4* b; float4* c; };     // (p, t), (ns_ff, bits: material | backside << 30, or -1 = no hit), (ng_ff, -)
CRT_D void store_hit(const HitRecords& H, int i, const SurfaceHitDev& h) {
    H.a[i] = make_float4(h.p.x, h.p.y, h.p.z, h.t);
    H.b[i] = make_float4(h.ns_ff.x, h.ns_ff.y, h.ns_ff.z, __int_as_float(h.found ? (h.material | (h.backside ? 0x40000000 : 0)) : -1));
    H.c[i] = make_float4(h.ng_ff.x, h.ng_ff.y, h.ng_ff.z, 0.f);
}
CRT_D void load_hit(const HitRecords& H, int i, SurfaceHitDev& h) {
    const float4 a = H.a[i], b = H.b[i], c = H.c[i];
    const int w = __float_as_int(b.w);
    h.found = w >= 0; h.material = w & 0x3fffffff; h.backside = (w >> 30) & 1;
    h.kind = -1; h.id0 = -1; h.id1 = -1;
    h.p = mk3(a.x, a.y, a.z); h.t = a.w; h.ns_ff = mk3(b.x, b.y, b.z); h.ng_ff = mk3(c.x, c.y, c.z);
}
struct MaterialQueues { int* ids; int* count; int capacity; };      // ids[type * capacity + slot], count[type]
#define CRT_STAGED_THREADS 128
#ifndef CRT_STAGED_MINB_HIT
#define CRT_STAGED_MINB_HIT 8          // CTAs per SM the compiler must fit: surface-record kernel, Lambert (light sample + BSDF), specular types
#endif
#ifndef CRT_STAGED_MINB_LAMBERT
#define CRT_STAGED_MINB_LAMBERT 6
#endif
#ifndef CRT_STAGED_MINB_SPECULAR
#define CRT_STAGED_MINB_SPECULAR 6
#endif
#ifndef CRT_STAGED_GRID
#define CRT_STAGED_GRID 32             // CTAs per SM of the persistent grids (C3 at 8 / 16 / 32: 593 / 632 / 646 Mpaths/s)
#endif

// closest_over_shapes for the 32 rays of a warp at once.  One ray per lane, a shape test costs ~10 box tests and a ray makes 1.3 of them
// after ~14 boxes: tested where they turn up, the lanes of a warp are at their shape tests at different times (ncu: 5.9 of 32 lanes per
// call; waiting for each other at the test instead makes the box walks wait, 27 % of lanes).  Here a lane only WALKS the hierarchy and
// drops every shape whose box it enters into a pool in shared memory; the warp empties the pool 32 (ray, shape) pairs at a time, each lane
// testing one pair with the owner's ray fetched by shuffle.  The answer is the same: a shape's BasicIntersect(ray, tMax) is its first
// valid root truncated at tMax, so the list-order loop returns the lexicographically smallest (t, index) among the shapes nearer than the
// mesh hit, and that minimum -- folded with a 64-bit atomicMin on (bits(t), index), t >= 0 -- does not depend on the order of the tests
// or on how far a lane's culling bound lags behind (a stale bound only lets more shapes into the pool).
#define CRT_HIT_POOL 96
struct HitWarpShared {
    unsigned long long key[32];        // per lane: (bits(t) << 32 | shape index) of the best shape so far; starts at (bits(mesh t or FLT_MAX), 0)
    float pay[6][32];                  // its object-space hit point and direction
    int pool[CRT_HIT_POOL];            // pending (owner lane << 16 | shape index)
    int n;
};
CRT_D void closest_over_shapes_warp(const DeviceScene& S, HitWarpShared& W, int lane, bool live, f3 ro, f3 rd, SurfaceHitDev& h) {
    float tMax = h.found ? h.t : FLT_MAX;
    const unsigned long long key0 = (unsigned long long)__float_as_uint(tMax) << 32;      // a shape must be strictly nearer than the mesh hit
    W.key[lane] = key0;
    if (lane == 0) W.n = 0;
    __syncwarp();
    RayConst rb;
    rb.o = ro; rb.inv_d = mk3(1 / rd.x, 1 / rd.y, 1 / rd.z);
    int i = live ? 0 : S.n_shape_nodes;
    while (true) {
        // walk: every lane on its own, until the hierarchy is exhausted or the pool holds two full batches
        while (i < S.n_shape_nodes && *(volatile int*)&W.n < CRT_HIT_POOL - 32) {
            const float4 lo = __ldg(&S.shape_bvh[2 * i]), hi = __ldg(&S.shape_bvh[2 * i + 1]);
            float m;
            if (!slab_unbounded(rb, lo, hi, m) || m > tMax) { i = __float_as_int(lo.w); continue; }      // cannot be hit within tMax: skip the subtree
            ++i;
            const int s = __float_as_int(hi.w);
            if (s >= 0) W.pool[atomicAdd(&W.n, 1)] = (lane << 16) | s;
        }
        __syncwarp();
        const bool all_done = __all_sync(CRT_FULL, i >= S.n_shape_nodes);
        const int n_pool = W.n;
        int pos = 0;
        while (n_pool - pos >= 32 || (all_done && pos < n_pool)) {
            const bool have = pos + lane < n_pool;
            const int e = have ? W.pool[pos + lane] : 0;
            const int owner = e >> 16, sidx = e & 0xffff;
            const f3 o = mk3(__shfl_sync(CRT_FULL, ro.x, owner), __shfl_sync(CRT_FULL, ro.y, owner), __shfl_sync(CRT_FULL, ro.z, owner));
            const f3 d = mk3(__shfl_sync(CRT_FULL, rd.x, owner), __shfl_sync(CRT_FULL, rd.y, owner), __shfl_sync(CRT_FULL, rd.z, owner));
            const unsigned long long cur = W.key[owner];
            unsigned long long key = ~0ull;
            ShapeIsect is;
            bool hit = false;
            // a root equal to the owner's bound still wins if its shape comes earlier in the list: test against the next float up
            if (have && shape_basic_lean(S.shapes[sidx], o, d, nextafterf(__uint_as_float((unsigned)(cur >> 32)), INFINITY), is) && is.t >= 0) {
                key = ((unsigned long long)__float_as_uint(is.t + 0.0f) << 32) | (unsigned)sidx;
                hit = key < cur;
                if (hit) atomicMin(&W.key[owner], key);
            }
            __syncwarp();
            if (hit && W.key[owner] == key) {
                W.pay[0][owner] = is.hitp.x; W.pay[1][owner] = is.hitp.y; W.pay[2][owner] = is.hitp.z;
                W.pay[3][owner] = is.ray_d.x; W.pay[4][owner] = is.ray_d.y; W.pay[5][owner] = is.ray_d.z;
            }
            __syncwarp();
            pos += 32;
        }
        const int rem = n_pool - pos;                  // < 32 pairs wait for the next round (none once every lane has finished its walk)
        const int keep = (lane < rem) ? W.pool[pos + lane] : 0;
        __syncwarp();
        if (lane < rem) W.pool[lane] = keep;
        if (lane == 0) W.n = rem > 0 ? rem : 0;
        __syncwarp();
        tMax = __uint_as_float((unsigned)(W.key[lane] >> 32));
        if (all_done) break;
    }
    const unsigned long long k = W.key[lane];
    if (!live || k >= key0) return;
    const int best = (int)(unsigned)(k & 0xffffffffu);
    const f3 best_p = mk3(W.pay[0][lane], W.pay[1][lane], W.pay[2][lane]), best_d = mk3(W.pay[3][lane], W.pay[4][lane], W.pay[5][lane]);
    int flipped;
    shape_surface_lean(S.shapes[best], best_p, best_d, h.p, h.ns_ff, flipped);
    h.found = 1; h.kind = 1; h.id0 = best; h.id1 = -1; h.t = tMax;
    h.ng_ff = h.ns_ff;
    h.backside = flipped;
    h.material = S.shapes[best].material;
}

__global__ void __launch_bounds__(CRT_STAGED_THREADS, CRT_STAGED_MINB_HIT) k_path_hit(DeviceScene S, PathBuffers pb, PathQueues Q, HitRecords H, MaterialQueues M) {
    __shared__ HitWarpShared shared[CRT_STAGED_THREADS / 32];
    HitWarpShared& W = shared[threadIdx.x >> 5];
    const int n = Q.n_active ? *Q.n_active : Q.n;
    const int lane = threadIdx.x & 31, warps = gridDim.x * (CRT_STAGED_THREADS / 32);
    for (int base = (blockIdx.x * (CRT_STAGED_THREADS / 32) + (threadIdx.x >> 5)) * 32; base < n; base += warps * 32) {
        const int slot = base + lane;
        const bool live = slot < n;
        int type = -1, i = 0;
        f3 ro = mk3(0, 0, 0), rd = mk3(0, 0, 1);
        SurfaceHitDev h;
        h.found = 0; h.kind = -1; h.id0 = -1; h.id1 = -1; h.material = 0; h.backside = 0; h.t = 0;
        h.p = mk3(0, 0, 0); h.ng_ff = h.p; h.ns_ff = h.p;
        if (live) {
            i = Q.active ? Q.active[slot] : slot;
            const float4 o4 = pb.ray_o[i], d4 = pb.ray_d[i];
            ro = mk3(o4.x, o4.y, o4.z); rd = mk3(d4.x, d4.y, d4.z);
            mesh_surface_hit(S, pb, i, ro, rd, o4.w, h);
        }
        if (S.n_shapes > 0) closest_over_shapes_warp(S, W, lane, live, ro, rd, h);
        if (live) {
            store_hit(H, i, h);
            if (h.found) type = S.materials[h.material].type;
        }
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const int q = warp_enqueue(type == t, M.count + t);
            if (type == t) M.ids[(size_t)t * M.capacity + q] = i;
        }
    }
}

template <int MAT>
__global__ void __launch_bounds__(CRT_STAGED_THREADS, MAT == MAT_LAMBERT ? CRT_STAGED_MINB_LAMBERT : CRT_STAGED_MINB_SPECULAR) k_path_shade_mat(DeviceScene S, RenderConst rc, PathBuffers pb, PathQueues Q, HitRecords H, MaterialQueues M) {
    const int n = M.count[MAT];
    const int* ids = M.ids + (size_t)MAT * M.capacity;
    const int lane = threadIdx.x & 31, warps = gridDim.x * (CRT_STAGED_THREADS / 32);
    for (int base = (blockIdx.x * (CRT_STAGED_THREADS / 32) + (threadIdx.x >> 5)) * 32; base < n; base += warps * 32) {
        const int slot = base + lane;
        int i = 0;
        ShadeOut out;
        out.continues = false; out.has_shadow = false; out.sh_o = make_float4(0, 0, 0, 0); out.sh_d = out.sh_o;
        if (slot < n) {
            i = ids[slot];
            const float4 d4 = pb.ray_d[i];
            SurfaceHitDev h;
            load_hit(H, i, h);
            shade_bounce<MAT>(S, rc, pb, i, mk3(d4.x, d4.y, d4.z), h, out);
        }
        enqueue_bounce(Q, i, out, MAT == MAT_LAMBERT);
    }
}

// Additional next-event slots of a bounce, one launch per light, BEFORE k_path_shade (they read the path state of the hit, which the
// shade kernel then advances): slot kind 0 = emissive triangle `index` sampled with probability 1 ("1 sample from each light source",
// Shading.h:4; draws one Get2D per light, in list order, ahead of the BSDF sample exactly like the oracle's loop), 1 = point light,
// 2 = sun (Lights.h:5-8; no random numbers).  Each slot has its own shadow queue; the queues are traced and added to L after the shade
// kernel, in slot order, which is the order of the oracle's `L +=`.
__global__ void __launch_bounds__(128) k_path_nee_slot(DeviceScene S, RenderConst rc, PathBuffers pb, PathQueues Q, HitRecords H, int slot_kind, int index) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = Q.n_active ? *Q.n_active : Q.n;
    LightSample ls;
    ls.valid = false; ls.o = make_float4(0, 0, 0, 0); ls.d = ls.o;
    int i = 0;
    if (slot < n) {
        i = Q.active ? Q.active[slot] : slot;
        const float4 o4 = pb.ray_o[i], d4 = pb.ray_d[i];
        const f3 ro = mk3(o4.x, o4.y, o4.z), rd = mk3(d4.x, d4.y, d4.z);
        SurfaceHitDev h;
        if (H.a) load_hit(H, i, h);                     // staged shading: k_path_hit has already formed the record
        else path_surface_hit(S, pb, i, ro, rd, o4.w, h);
        const int depth = (int)(((unsigned)pb.flags[i]) >> 8);
        if (h.found && depth != rc.max_depth) {
            const DevMaterial m = S.materials[h.material];
            if (m.type == MAT_LAMBERT && m.refl >= 0) {
                Spec8 lambda, beta, R;
                load8(pb.lambda, i, lambda); load8(pb.beta, i, beta);
                spectrum_sample(S, m.refl, lambda, R);
                if (slot_kind == 0) {
                    SamplerState ss = pb.sampler[i];
                    const f2 up = sampler_get2d(rc.sampler, ss);
                    pb.sampler[i] = ss;
                    sample_emissive_triangle(S, rc, h, lambda, beta, R, S.lights[index], 1.0f, up, ls);
                } else {
                    const DevDeltaLight dl = S.delta_lights[index];
                    f3 so, wl;
                    float tmax, atten;
                    if (dl.kind == 0) {
                        const f3 lp = mk3(dl.v[0], dl.v[1], dl.v[2]);
                        so = offset_origin(h.p, h.ng_ff, lp - h.p, rc.ray_eps);
                        const f3 dvec = lp - so;
                        const float dist2 = dot3(dvec, dvec);
                        const float dist = sqrtf(dist2);
                        wl = dvec * (1.0f / dist);
                        tmax = dist * (1 - rc.shadow_eps);
                        atten = 1.0f / dist2;
                    } else {
                        wl = mk3(dl.v[0], dl.v[1], dl.v[2]);
                        so = offset_origin(h.p, h.ng_ff, wl, rc.ray_eps);
                        tmax = FLT_MAX;
                        atten = 1.0f;
                    }
                    const float cos_s = dot3(h.ns_ff, wl);
                    if (cos_s > 0 && dot3(h.ng_ff, wl) > 0) {
                        const float g = cos_s * atten;
#pragma unroll
                        for (int k = 0; k < CRT_NLAMBDA; ++k) {
                            const float I = spectrum_query(S, dl.spectrum, lambda.v[k]) * dl.scale;
                            ls.contrib.v[k] = ((beta.v[k] * (R.v[k] * CRT_INV_PI)) * I) * g;
                        }
                        ls.valid = true;
                        ls.o = make_float4(so.x, so.y, so.z, tmax);
                        ls.d = make_float4(wl.x, wl.y, wl.z, 0);
                    }
                }
            }
        }
    }
    const int s_slot = warp_enqueue(ls.valid, Q.n_shadow);
    if (ls.valid) {
        store_ray(Q.sh_o, Q.sh_d, Q.sh_k, Q.sh_s, s_slot, mk3(ls.o.x, ls.o.y, ls.o.z), mk3(ls.d.x, ls.d.y, ls.d.z), ls.o.w);
        Q.sh_path[s_slot] = i;
        store8(Q.sh_contrib, s_slot, ls.contrib);
    }
}

// occluded_by_shapes for the 32 shadow rays of a warp, pooled like closest_over_shapes_warp: lanes walk the hierarchy and drop (ray, shape)
// pairs into the pool, the warp tests 32 pairs at a time; the first hit of a ray ends its walk.  tMax is fixed, so the answer (any shape
// hit within tMax) does not depend on the order of the tests.  (Per lane this loop ran at 14 of 32 lanes on C3.)
CRT_D bool occluded_by_shapes_warp(const DeviceScene& S, HitWarpShared& W, int lane, bool live, f3 ro, f3 rd, float tMax) {
    volatile unsigned long long* occ = W.key;          // per lane: 1 = some shape occludes this lane's ray
    occ[lane] = 0;
    if (lane == 0) W.n = 0;
    __syncwarp();
    RayConst rb;
    rb.o = ro; rb.inv_d = mk3(1 / rd.x, 1 / rd.y, 1 / rd.z);
    int i = live ? 0 : S.n_shape_nodes;
    while (true) {
        while (i < S.n_shape_nodes && *(volatile int*)&W.n < CRT_HIT_POOL - 32) {
            const float4 lo = __ldg(&S.shape_bvh[2 * i]), hi = __ldg(&S.shape_bvh[2 * i + 1]);
            float m;
            if (!slab_unbounded(rb, lo, hi, m) || m > tMax) { i = __float_as_int(lo.w); continue; }
            ++i;
            const int s = __float_as_int(hi.w);
            if (s >= 0) W.pool[atomicAdd(&W.n, 1)] = (lane << 16) | s;
        }
        __syncwarp();
        const bool all_done = __all_sync(CRT_FULL, i >= S.n_shape_nodes);
        const int n_pool = W.n;
        int pos = 0;
        while (n_pool - pos >= 32 || (all_done && pos < n_pool)) {
            const bool have = pos + lane < n_pool;
            const int e = have ? W.pool[pos + lane] : 0;
            const int owner = e >> 16, sidx = e & 0xffff;
            const f3 o = mk3(__shfl_sync(CRT_FULL, ro.x, owner), __shfl_sync(CRT_FULL, ro.y, owner), __shfl_sync(CRT_FULL, ro.z, owner));
            const f3 d = mk3(__shfl_sync(CRT_FULL, rd.x, owner), __shfl_sync(CRT_FULL, rd.y, owner), __shfl_sync(CRT_FULL, rd.z, owner));
            const float tm = __shfl_sync(CRT_FULL, tMax, owner);
            ShapeIsect is;
            if (have && occ[owner] == 0 && shape_basic_lean(S.shapes[sidx], o, d, tm, is)) occ[owner] = 1;
            __syncwarp();
            pos += 32;
        }
        const int rem = n_pool - pos;
        const int keep = (lane < rem) ? W.pool[pos + lane] : 0;
        __syncwarp();
        if (lane < rem) W.pool[lane] = keep;
        if (lane == 0) W.n = rem > 0 ? rem : 0;
        __syncwarp();
        if (occ[lane] != 0) i = S.n_shape_nodes;          // decided: stop walking (pairs of this ray still in the pool are skipped)
        if (all_done) break;
    }
    return occ[lane] != 0;
}

// L[path] += contribution of every shadow ray that reached its light (oracle_render.cpp:229-234)
__global__ void __launch_bounds__(256) k_shadow_resolve(DeviceScene S, PathBuffers pb, PathQueues Q, const int* occluded) {
    __shared__ HitWarpShared shared[256 / 32];
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0) {           // bookkeeping between bounces: this bounce's ray counts are final by now
        if (Q.count_active) Q.ray_counters[0] += (unsigned long long)(Q.n_active ? *Q.n_active : Q.n);
        Q.ray_counters[1] += (unsigned long long)*Q.n_shadow;
    }
    const int n = *Q.n_shadow;
    if ((s & ~31) >= n) return;                          // the whole warp is past the queue
    const bool live = s < n;
    bool occ = false;
    f3 so = mk3(0, 0, 0), sd = mk3(0, 0, 1);
    float tmax = 0;
    if (live) {
        occ = (S.has_model && !S.root_leaf) ? occluded[s] != 0 : false;
        if (S.root_leaf || S.n_shapes > 0) {
            const float4 o4 = Q.sh_o[s], d4 = Q.sh_d[s];
            so = mk3(o4.x, o4.y, o4.z); sd = mk3(d4.x, d4.y, d4.z); tmax = o4.w;
            if (S.root_leaf) { int ref; float4 tb; occ = trace_root_leaf<true>(S, so, sd, tmax, ref, tb); }
        }
    }
    if (S.n_shapes > 0) {
        const bool want = live && !occ;
        const bool hit = occluded_by_shapes_warp(S, shared[threadIdx.x >> 5], threadIdx.x & 31, want, so, sd, tmax);
        if (want) occ = hit;
    }
    if (!live || occ) return;
    int i = Q.sh_path[s];           // at most one shadow ray per path per bounce: no race on L[i]
    Spec8 L, c;
    load8(pb.L, i, L); load8(Q.sh_contrib, s, c);
#pragma unroll
    for (int k = 0; k < CRT_NLAMBDA; ++k) L.v[k] += c.v[k];
    store8(pb.L, i, L);
}

// the same bookkeeping for scenes without lights (no k_shadow_resolve launch): accumulate ray counts (runs as one thread)
__global__ void k_path_count(PathQueues Q, int bounce) {
    int n = Q.n_active ? *Q.n_active : Q.n;
    Q.ray_counters[0] += (unsigned long long)n;
    Q.ray_counters[1] += (unsigned long long)*Q.n_shadow;
    (void)bounce;
}

// end of the wave: ToSensorRGB + clamp + film accumulate for every path slot (RayTracerTestApp.h:326-337)
__global__ void __launch_bounds__(256) k_path_splat(DeviceScene S, PathBuffers pb, float4* film, SampleDebugOut dbg, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;

    Spec8 lambda, pdf, L;
    load8(pb.lambda, i, lambda); load8(pb.pdf, i, pdf); load8(pb.L, i, L);
    splat_or_debug(S, pb, i, L, lambda, pdf, film, dbg);
}
// probe: the camera rays of a wave, before the bounces overwrite them
__global__ void k_dump_rays(PathBuffers pb, float* ray6, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 o4 = pb.ray_o[i], d4 = pb.ray_d[i];
    ray6[6 * i] = o4.x; ray6[6 * i + 1] = o4.y; ray6[6 * i + 2] = o4.z;
    ray6[6 * i + 3] = d4.x; ray6[6 * i + 4] = d4.y; ray6[6 * i + 5] = d4.z;
}
// The same for a wave that holds `nsamp` consecutive sample indices of every pixel (slot = s * n_pix + p): one thread per
// pixel slot adds its samples in ascending index order, i.e. exactly the sequence of `+=` the one-index-per-wave loop
// (and the reference's pixel_index loop, RayTracerTestApp.h:399-422) performs on that pixel.
__global__ void __launch_bounds__(256) k_path_splat_multi(DeviceScene S, PathBuffers pb, float4* film, int n_pix, int nsamp) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned depth = 0;
    if (p < n_pix) {
    const int pixel = pb.pixel[p];
    float4 f = film[pixel];
    for (int s = 0; s < nsamp; ++s) {
        const size_t i = (size_t)s * n_pix + p;
        depth += ((unsigned)pb.flags[i]) >> 8;
        Spec8 lambda, pdf, L;
        load8(pb.lambda, i, lambda); load8(pb.pdf, i, pdf); load8(pb.L, i, L);
        f3 cam = to_sensor_rgb(S, L, lambda, pdf);
        cam.x = gclamp(cam.x, 0.0f, 1.0f); cam.y = gclamp(cam.y, 0.0f, 1.0f); cam.z = gclamp(cam.z, 0.0f, 1.0f);   // RayTracerTestApp.h:332-334
        const float w = pb.weight[i];
        f.x += w * cam.x; f.y += w * cam.y; f.z += w * cam.z; f.w += w;
    }
    film[pixel] = f;
    }
    if (pb.depth_sum) {          // realised path depths (statistics), one atomic per warp
        depth = __reduce_add_sync(CRT_FULL, depth);
        if ((threadIdx.x & 31) == 0 && depth) atomicAdd(pb.depth_sum, (unsigned long long)depth);
    }
}
__global__ void __launch_bounds__(256) k_path_depth_sum(PathBuffers pb, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned d = i < n ? (((unsigned)pb.flags[i]) >> 8) : 0u;
    d = __reduce_add_sync(CRT_FULL, d);
    if ((threadIdx.x & 31) == 0 && d) atomicAdd(pb.depth_sum, (unsigned long long)d);
}

}  // namespace crt
