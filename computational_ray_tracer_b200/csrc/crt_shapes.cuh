// crt_shapes.cuh -- analytic shapes on the device: Sphere, Cylinder, Disk, TriangleSimple
// (RayTracer/Shapes.h:209-905).  One thread evaluates one (ray, shape) pair; the operation order is the
// reference's.  atan2f / acosf / sinf / cosf are CUDA's (<= 2 ulp from glibc's): quantities derived from
// them (phi clipping at partial sweeps, u, v, dv) are tolerance-checked, everything else is bit-exact.
#pragma once
#include "crt_device_scene.h"

namespace crt {

struct ShapeIsect {
    float t;
    f3 hitp;        // object space
    f3 ray_d;       // object-space direction as the reference stores it (normalised except for Cylinder)
    float phi;
    float B, Y;     // TriangleSimple barycentrics
};
struct SurfaceInfo {           // LocalSurfaceInfo (Shapes.h:144-170), world space
    float tHit;
    f3 hitp, n;
    float u, v;
    f3 du, dv, wo;
    int flipped;    // the face-forward test reversed the outward normal
};

CRT_D float wrap_phi(float y, float x) {
    float phi = atan2f(y, x);
    if (phi < 0) phi = (float)((double)phi + 2 * 3.141592653589793238462643383279502884);   // float += double, Shapes.h:317
    return phi;
}

// The integrator needs phi only for the partial-sweep test `phi > phimax` (PHI = false; the probes also report it, PHI = true).  wrap_phi
// never exceeds float(2 pi) -- atan2f is in [-pi, pi] and a negative value plus the double 2 pi rounds to at most 6.2831855f -- so for a
// complete sweep (phimax = 360 deg * pi/180 = 6.2831855f) the test cannot hold and atan2f is skipped.
#define CRT_FULL_SWEEP 6.28318548f
template <bool PHI> CRT_D float clip_phi(float y, float x, float phimax) {
    if (!PHI && phimax >= CRT_FULL_SWEEP) return 0.f;
    return wrap_phi(y, x);
}

CRT_D bool quadratic_roots(float a, float b, float c, float r, float len, float tMax, float& t0, float& t1, float& tHit) {
    float discrim = 4 * a * (r + len) * (r - len);
    if (discrim < 0) return false;
    float rootDiscrim = sqrtf(discrim);
    float q = (b < 0) ? -.5f * (b - rootDiscrim) : -.5f * (b + rootDiscrim);
    t0 = q / a; t1 = c / q;
    if (t0 > t1) { float tmp = t1; t1 = t0; t0 = tmp; }
    if (t0 > tMax || t1 <= 0) return false;
    tHit = t0;
    if (tHit <= 0) {
        tHit = t1;
        if (tHit > tMax) return false;
    }
    return true;
}

// Sphere::BasicIntersect, Shapes.h:277-357.  p = {r, zmin, zmax, thetamin, thetamax, phimax}
template <bool PHI> CRT_D bool sphere_basic(const DevShape& s, f3 ro, f3 rd, float tMax, ShapeIsect& is) {
    const float r = s.p[0], zmin = s.p[1], zmax = s.p[2], phimax = s.p[5];
    f3 o = xform_point(s.r2o, ro), d = xform_vector(s.r2o, rd);
    float a = d.x * d.x + d.y * d.y + d.z * d.z;
    float b = 2 * (d.x * o.x + d.y * o.y + d.z * o.z);
    float c = o.x * o.x + o.y * o.y + o.z * o.z - r * r;
    f3 v = o - (b / (2 * a)) * d;
    float len = length3(v);
    float t0, t1, tHit;
    if (!quadratic_roots(a, b, c, r, len, tMax, t0, t1, tHit)) return false;
    f3 hitp; float phi;
    for (int attempt = 0; attempt < 2; ++attempt) {
        hitp = o + tHit * d;
        hitp = hitp * (r / length3(mk3(0, 0, 0) - hitp));        // glm::distance(hitp, 0) = length(0 - hitp)
        if (hitp.x == 0 && hitp.y == 0) hitp.x = (float)(1e-5 * (double)r);
        phi = clip_phi<PHI>(hitp.y, hitp.x, phimax);
        if (!(hitp.z < zmin || hitp.z > zmax || phi > phimax)) break;
        if (attempt == 1) return false;
        if (tHit == t1) return false;
        if (t1 > tMax) return false;
        tHit = t1;
    }
    is.t = tHit; is.hitp = hitp; is.ray_d = normalize3(d); is.phi = phi;
    return true;
}
// Cylinder::BasicIntersect, Shapes.h:499-564.  p = {r, min_z, max_z, max_phi}
template <bool PHI> CRT_D bool cylinder_basic(const DevShape& s, f3 ro, f3 rd, float tMax, ShapeIsect& is) {
    const float r = s.p[0], zmin = s.p[1], zmax = s.p[2], phimax = s.p[3];
    f3 o = xform_point(s.r2o, ro), d = xform_vector(s.r2o, rd);
    float a = d.x * d.x + d.y * d.y;
    float b = 2 * (d.x * o.x + d.y * o.y);
    float c = o.x * o.x + o.y * o.y - r * r;
    float f = b / (2 * a);
    float vx = o.x - f * d.x, vy = o.y - f * d.y;
    float len = sqrtf(vx * vx + vy * vy);
    float t0, t1, tHit;
    if (!quadratic_roots(a, b, c, r, len, tMax, t0, t1, tHit)) return false;
    f3 hp = o + tHit * d;
    float phi = clip_phi<PHI>(hp.y, hp.x, phimax);
    if (hp.z < zmin || hp.z > zmax || phi > phimax) {
        if (tHit == t1) return false;
        tHit = t1;
        if (t1 > tMax) return false;
        hp = o + tHit * d;
        phi = clip_phi<PHI>(hp.y, hp.x, phimax);
        if (hp.z < zmin || hp.z > zmax || phi > phimax) return false;
    }
    is.t = tHit; is.hitp = hp; is.ray_d = d; is.phi = phi;      // NB: the cylinder keeps d un-normalised (:563)
    return true;
}
// Disk::BasicIntersect, Shapes.h:684-710.  p = {h, inner_r, outer_r, phimax}
template <bool PHI> CRT_D bool disk_basic(const DevShape& s, f3 ro, f3 rd, float tMax, ShapeIsect& is) {
    const float h = s.p[0], inner = s.p[1], outer = s.p[2], phimax = s.p[3];
    f3 o = xform_point(s.r2o, ro), d = xform_vector(s.r2o, rd);
    float t0 = (h - o.z) / d.z;
    if (t0 <= 0 || t0 >= tMax) return false;
    if (d.z == 0) return false;
    f3 ph = o + t0 * d;
    float dist2 = ph.x * ph.x + ph.y * ph.y;
    if (dist2 > outer * outer || dist2 < inner * inner) return false;
    float phi = clip_phi<PHI>(ph.y, ph.x, phimax);
    if (phi > phimax) return false;
    is.t = t0; is.hitp = ph; is.ray_d = normalize3(d); is.phi = phi;
    return true;
}
// TriangleSimple::BasicIntersect, Shapes.h:830-869 (Cramer's rule).  p = {p1, p2, p3}
CRT_D bool trisimple_basic(const DevShape& s, f3 ro, f3 rd, float tMax, ShapeIsect& is) {
    f3 orig = xform_point(s.r2o, ro), dir = xform_vector(s.r2o, rd);
    const float* P = s.p;
    float a = P[0] - P[3], b = P[1] - P[4], c = P[2] - P[5];
    float d = P[0] - P[6], e = P[1] - P[7], f = P[2] - P[8];
    float g = dir.x, h = dir.y, i = dir.z;
    float j = P[0] - orig.x, k = P[1] - orig.y, l = P[2] - orig.z;
    float M = a * (e * i - h * f) + b * (g * f - d * i) + c * (d * h - e * g);
    float t = -(f * (a * k - j * b) + e * (j * c - a * l) + d * (b * l - k * c)) / M;
    if (t < 0 || t >= tMax) return false;
    float Y = (i * (a * k - j * b) + h * (j * c - a * l) + g * (b * l - k * c)) / M;
    if (Y < 0 || Y > 1) return false;
    float B = (j * (e * i - h * f) + k * (g * f - d * i) + l * (d * h - e * g)) / M;
    if (B < 0 || B > 1 - Y) return false;
    is.t = t; is.hitp = orig + t * dir; is.ray_d = normalize3(dir); is.B = B; is.Y = Y; is.phi = 0;
    return true;
}
// (not inlined: see spectrum_query_nl for why code size matters in the integrator)
__device__ __noinline__ bool shape_basic(const DevShape& s, f3 ro, f3 rd, float tMax, ShapeIsect& is) {
    switch (s.kind) {
        case SHAPE_SPHERE: return sphere_basic<true>(s, ro, rd, tMax, is);
        case SHAPE_CYLINDER: return cylinder_basic<true>(s, ro, rd, tMax, is);
        case SHAPE_DISK: return disk_basic<true>(s, ro, rd, tMax, is);
        default: return trisimple_basic(s, ro, rd, tMax, is);
    }
}
// the integrator's version: same hit / miss, t, hitp and ray_d; is.phi is not meaningful
__device__ __noinline__ bool shape_basic_lean(const DevShape& s, f3 ro, f3 rd, float tMax, ShapeIsect& is) {
    switch (s.kind) {
        case SHAPE_SPHERE: return sphere_basic<false>(s, ro, rd, tMax, is);
        case SHAPE_CYLINDER: return cylinder_basic<false>(s, ro, rd, tMax, is);
        case SHAPE_DISK: return disk_basic<false>(s, ro, rd, tMax, is);
        default: return trisimple_basic(s, ro, rd, tMax, is);
    }
}
// Shape::Intersect's surface record (Shapes.h:244-270, 465-492, 653-679, 797-824) + LocalSurfaceInfo::Transform
__device__ __noinline__ void shape_surface(const DevShape& s, const ShapeIsect& is, SurfaceInfo& out) {
    f3 p = is.hitp, n, du, dv;
    float u, v;
    if (s.kind == SHAPE_SPHERE) {
        const float r = s.p[0], thetamin = s.p[3], thetamax = s.p[4], phimax = s.p[5];
        float theta = acosf(gclamp(p.z / r, -1.f, 1.f));
        float phi = wrap_phi(p.y, p.x);
        u = phi / phimax; v = (theta - thetamin) / (thetamax - thetamin);
        du = normalize3(mk3(-phimax * p.y, phimax * p.x, 0));                                                   // calculate_du, Shapes.h:393-396
        dv = normalize3((thetamax - thetamin) * mk3(p.z * cosf(phi), p.z * sinf(phi), -r * sinf(theta)));        // calculate_dv, :401-408
        n = normalize3(mk3(2 * p.x, 2 * p.y, 2 * p.z));
    } else if (s.kind == SHAPE_CYLINDER) {
        const float zmin = s.p[1], zmax = s.p[2], phimax = s.p[3];
        float phi = wrap_phi(p.y, p.x);
        u = phi / phimax; v = (p.z - zmin) / (zmax - zmin);
        du = normalize3(mk3(-phimax * p.y, phimax * p.x, 0));                                                   // :589-592
        dv = normalize3(mk3(0, 0, zmax - zmin));                                                                // :597-600
        n = normalize3(mk3(2 * p.x, 2 * p.y, 0));
    } else if (s.kind == SHAPE_DISK) {
        const float inner = s.p[1], outer = s.p[2], phimax = s.p[3];
        float phi = wrap_phi(p.y, p.x);
        const float rad = sqrtf(p.x * p.x + p.y * p.y);
        u = phi / phimax; v = (outer - rad) / (outer - inner);
        du = normalize3(mk3(-phimax * p.y, phimax * p.x, 0));                                                   // :731-734
        dv = normalize3((mk3(p.x, p.y, 0) * (inner - outer)) / rad);                                            // :736-739
        n = mk3(0, 0, 1);
    } else {
        const float* P = s.p;
        f3 p1 = mk3(P[0], P[1], P[2]), p2 = mk3(P[3], P[4], P[5]), p3 = mk3(P[6], P[7], P[8]);
        u = is.B; v = is.Y;
        du = normalize3(p2 - p1);                                                                               // :884-887
        dv = normalize3(p3 - p1);                                                                               // :889-892
        n = normalize3(cross3(p3 - p1, p2 - p1));
    }
    out.flipped = dot3(n, is.ray_d) > 0;
    if (out.flipped) n = -n;
    out.tHit = is.t;
    out.u = gclamp(u, 0.0f, 1.0f);
    out.v = gclamp(v, 0.0f, 1.0f);
    // LocalSurfaceInfo::Transform(ObjectToRender) (Shapes.h:147-160)
    out.n = normalize3(mul_m3_v3(s.nmat, n));
    out.wo = normalize3(mul_m3_v3(s.nmat, mk3(0, 0, 1)));
    out.hitp = xform_point(s.o2r, p);
    out.du = xform_vector(s.o2r, du);
    out.dv = xform_vector(s.o2r, dv);
}
// The part of that record the path integrator uses -- position, face-forwarded normal, which side was hit -- by the operations of
// shape_surface that produce them (u, v, du, dv, wo and their acosf / atan2f / sinf / cosf are left out).
__device__ __noinline__ void shape_surface_lean(const DevShape& s, f3 p, f3 ray_d, f3& hitp, f3& n_out, int& flipped) {
    f3 n;
    if (s.kind == SHAPE_SPHERE) n = normalize3(mk3(2 * p.x, 2 * p.y, 2 * p.z));
    else if (s.kind == SHAPE_CYLINDER) n = normalize3(mk3(2 * p.x, 2 * p.y, 0));
    else if (s.kind == SHAPE_DISK) n = mk3(0, 0, 1);
    else {
        const float* P = s.p;
        f3 p1 = mk3(P[0], P[1], P[2]), p2 = mk3(P[3], P[4], P[5]), p3 = mk3(P[6], P[7], P[8]);
        n = normalize3(cross3(p3 - p1, p2 - p1));
    }
    flipped = dot3(n, ray_d) > 0;
    if (flipped) n = -n;
    n_out = normalize3(mul_m3_v3(s.nmat, n));
    hitp = xform_point(s.o2r, p);
}

}  // namespace crt
