// crt_math.h -- fp32 vector helpers shared by the host builder and the sm_100a kernels.
//
// Every function fixes one evaluation order (no FMA contraction: the library is built with
// nvcc --fmad=false and g++ -ffp-contract=off) so host and device produce the same bits as the
// reference's glm-based expressions evaluated left to right.  fmaf appears only where the reference
// itself calls std::fma (ThirdParty/pbrv4/helpers.h:56-62).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define CRT_HD __host__ __device__ __forceinline__
#define CRT_D __device__ __forceinline__
#else
#define CRT_HD inline
#define CRT_D inline
#endif

namespace crt {

struct f2 { float x, y; };
struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };

CRT_HD f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
CRT_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
CRT_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
CRT_HD f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
CRT_HD f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
CRT_HD f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
CRT_HD f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
CRT_HD float comp(f3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

// glm::dot(vec3): three products, then (x + y) + z
CRT_HD float dot3(f3 a, f3 b) {
    float px = a.x * b.x, py = a.y * b.y, pz = a.z * b.z;
    return (px + py) + pz;
}
// glm::cross
CRT_HD f3 cross3(f3 a, f3 b) { return mk3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
CRT_HD float length3(f3 v) { return sqrtf(dot3(v, v)); }
// glm::normalize = v * (1 / sqrt(dot(v, v)))
CRT_HD f3 normalize3(f3 v) { float s = 1.0f / sqrtf(dot3(v, v)); return v * s; }
// glm::min / glm::max / glm::clamp (ternary forms keep the reference's NaN behaviour)
CRT_HD float gmin(float x, float y) { return (y < x) ? y : x; }
CRT_HD float gmax(float x, float y) { return (x < y) ? y : x; }
CRT_HD float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
// std::max({a,b,c}) == *max_element: keeps the first largest under operator<
CRT_HD float max3_std(float a, float b, float c) { float m = a; if (m < b) m = b; if (m < c) m = c; return m; }
CRT_HD float max_std(float a, float b) { return (a < b) ? b : a; }      // std::max
CRT_HD float min_std(float a, float b) { return (b < a) ? b : a; }      // std::min

// pbrt::gamma(n) (helpers.h:50-54), evaluated in fp32 exactly like the constexpr in the reference
CRT_HD constexpr float gamma_n(int n) { return (n * (FLT_EPSILON * 0.5f)) / (1 - n * (FLT_EPSILON * 0.5f)); }
// pbrt::DifferenceOfProducts (helpers.h:56-62)
CRT_HD float diff_of_products(float a, float b, float c, float d) {
    float cd = c * d;
    float dop = fmaf(a, b, -cd);
    float err = fmaf(-c, d, cd);
    return dop + err;
}
CRT_HD float lerp_pbrt(float x, float a, float b) { return (1 - x) * a + x * b; }      // helpers.h:154-157
CRT_HD float safe_sqrt(float x) { return sqrtf(max_std(0.f, x)); }                      // helpers.h:174-178

// column-major 4x4 (glm layout): m[col*4 + row]
struct m4 { float m[16]; };
// glm mat4 * vec4, scalar path: (c0*x + c1*y) + (c2*z + c3*w)
CRT_HD f4 mul_m4_v4(const float* m, float x, float y, float z, float w) {
    f4 r;
    r.x = (m[0] * x + m[4] * y) + (m[8] * z + m[12] * w);
    r.y = (m[1] * x + m[5] * y) + (m[9] * z + m[13] * w);
    r.z = (m[2] * x + m[6] * y) + (m[10] * z + m[14] * w);
    r.w = (m[3] * x + m[7] * y) + (m[11] * z + m[15] * w);
    return r;
}
CRT_HD f3 xform_point(const float* m, f3 p) { f4 r = mul_m4_v4(m, p.x, p.y, p.z, 1.0f); return mk3(r.x, r.y, r.z); }
CRT_HD f3 xform_vector(const float* m, f3 v) { f4 r = mul_m4_v4(m, v.x, v.y, v.z, 0.0f); return mk3(r.x, r.y, r.z); }
// glm::normalize(vec4) then .xyz -- Ray::Transform (Shapes.h:37-41) normalises the 4-vector (w = 0 contributes 0*0)
CRT_HD f3 xform_dir_normalized(const float* m, f3 v) {
    f4 r = mul_m4_v4(m, v.x, v.y, v.z, 0.0f);
    float px = r.x * r.x, py = r.y * r.y, pz = r.z * r.z, pw = r.w * r.w;
    float s = 1.0f / sqrtf((px + py) + (pz + pw));
    return mk3(r.x * s, r.y * s, r.z * s);
}
// glm mat3 (stored as 9 floats column-major) * vec3: (m00*x + m10*y) + m20*z
CRT_HD f3 mul_m3_v3(const float* m, f3 v) {
    return mk3((m[0] * v.x + m[3] * v.y) + m[6] * v.z, (m[1] * v.x + m[4] * v.y) + m[7] * v.z, (m[2] * v.x + m[5] * v.y) + m[8] * v.z);
}

}  // namespace crt
