// ORACLE SUPPORT (test infrastructure, never shipped, never on the product path).
//
// A small stand-in for the subset of g-truc/glm that the reference's render path uses
// (pch.h:25-30 includes glm; the reference neither vendors it nor pins a version).  It
// exists so that the reference's OWN headers (RayTracer/*.h, ThirdParty/pbrv4/*,
// ThirdParty/AABB_triangle_Moller.h) can be compiled unmodified, where they lie, into
// oracle/_ref/libcrt_ref.so (oracle/Makefile target `ref`) and run against the restated
// oracle.  Written from scratch; every function fixes the same evaluation order as
// oracle/ovec.h (glm's scalar code path), built with -ffp-contract=off.
//
// What this pins and what it does not: everything ABOVE the glm boundary (the reference's
// slab test, watertight triangle test, BFS traversal, octree build, SAT, samplers, spectra,
// sensor, cameras, analytic shapes) is the reference's code, bit for bit.  The order of
// operations INSIDE mat4*vec4, inverse, normalize, cross, dot is this file's choice, as it
// is ovec.h's (SURVEY.md 8c "glm boundary").
#pragma once
#include <cmath>
#include <cstddef>
#include <string>
#include <limits>
#include <type_traits>

namespace glm {

typedef int length_t;

template <typename T> struct tvec2 {
    union { T x, r, s; };
    union { T y, g, t; };
    constexpr tvec2() : x(0), y(0) {}
    constexpr tvec2(T a, T b) : x(a), y(b) {}
    constexpr explicit tvec2(T a) : x(a), y(a) {}
    template <typename A, typename B> constexpr tvec2(A a, B b) : x(T(a)), y(T(b)) {}
    template <typename U> constexpr tvec2(const tvec2<U>& v) : x(T(v.x)), y(T(v.y)) {}
    T& operator[](int i) { return i == 0 ? x : y; }
    constexpr const T& operator[](int i) const { return i == 0 ? x : y; }
    static constexpr length_t length() { return 2; }
};
template <typename T> struct tvec4;
template <typename T> struct tvec3 {
    union { T x, r, s; };
    union { T y, g, t; };
    union { T z, b, p; };
    constexpr tvec3() : x(0), y(0), z(0) {}
    constexpr tvec3(T a, T b_, T c) : x(a), y(b_), z(c) {}
    constexpr explicit tvec3(T a) : x(a), y(a), z(a) {}
    template <typename A, typename B, typename C> constexpr tvec3(A a, B b_, C c) : x(T(a)), y(T(b_)), z(T(c)) {}
    template <typename U> constexpr tvec3(const tvec3<U>& v) : x(T(v.x)), y(T(v.y)), z(T(v.z)) {}
    template <typename U> constexpr tvec3(const tvec4<U>& v);  // glm: implicit unless GLM_FORCE_EXPLICIT_CTOR
    template <typename U, typename C> constexpr tvec3(const tvec2<U>& v, C c) : x(T(v.x)), y(T(v.y)), z(T(c)) {}
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    constexpr const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    static constexpr length_t length() { return 3; }
};
template <typename T> struct tvec4 {
    union { T x, r, s; };
    union { T y, g, t; };
    union { T z, b, p; };
    union { T w, a, q; };
    constexpr tvec4() : x(0), y(0), z(0), w(0) {}
    constexpr tvec4(T a_, T b_, T c, T d) : x(a_), y(b_), z(c), w(d) {}
    constexpr explicit tvec4(T a_) : x(a_), y(a_), z(a_), w(a_) {}
    template <typename A, typename B, typename C, typename D>
    constexpr tvec4(A a_, B b_, C c, D d) : x(T(a_)), y(T(b_)), z(T(c)), w(T(d)) {}
    template <typename U, typename D> constexpr tvec4(const tvec3<U>& v, D d) : x(T(v.x)), y(T(v.y)), z(T(v.z)), w(T(d)) {}
    template <typename U> constexpr tvec4(const tvec4<U>& v) : x(T(v.x)), y(T(v.y)), z(T(v.z)), w(T(v.w)) {}
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    constexpr const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    static constexpr length_t length() { return 4; }
};
template <typename T> template <typename U>
constexpr tvec3<T>::tvec3(const tvec4<U>& v) : x(T(v.x)), y(T(v.y)), z(T(v.z)) {}

typedef tvec2<float> vec2;   typedef tvec3<float> vec3;   typedef tvec4<float> vec4;
typedef vec2 fvec2; typedef vec3 fvec3; typedef vec4 fvec4;
typedef tvec2<double> dvec2; typedef tvec3<double> dvec3; typedef tvec4<double> dvec4;
typedef tvec2<int> ivec2;    typedef tvec3<int> ivec3;    typedef tvec4<int> ivec4;
typedef tvec2<unsigned> uvec2; typedef tvec3<unsigned> uvec3; typedef tvec4<unsigned> uvec4;
typedef tvec2<bool> bvec2;   typedef tvec3<bool> bvec3;   typedef tvec4<bool> bvec4;

// ---- component-wise arithmetic -------------------------------------------------------------
#define CRT_GLM_BINOP(OP)                                                                               \
    template <typename T> constexpr tvec2<T> operator OP(const tvec2<T>& a, const tvec2<T>& b) { return {a.x OP b.x, a.y OP b.y}; } \
    template <typename T> constexpr tvec2<T> operator OP(const tvec2<T>& a, T s) { return {a.x OP s, a.y OP s}; }                   \
    template <typename T> constexpr tvec2<T> operator OP(T s, const tvec2<T>& a) { return {s OP a.x, s OP a.y}; }                   \
    template <typename T> constexpr tvec3<T> operator OP(const tvec3<T>& a, const tvec3<T>& b) { return {a.x OP b.x, a.y OP b.y, a.z OP b.z}; } \
    template <typename T> constexpr tvec3<T> operator OP(const tvec3<T>& a, T s) { return {a.x OP s, a.y OP s, a.z OP s}; }         \
    template <typename T> constexpr tvec3<T> operator OP(T s, const tvec3<T>& a) { return {s OP a.x, s OP a.y, s OP a.z}; }         \
    template <typename T> constexpr tvec4<T> operator OP(const tvec4<T>& a, const tvec4<T>& b) { return {a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w}; } \
    template <typename T> constexpr tvec4<T> operator OP(const tvec4<T>& a, T s) { return {a.x OP s, a.y OP s, a.z OP s, a.w OP s}; } \
    template <typename T> constexpr tvec4<T> operator OP(T s, const tvec4<T>& a) { return {s OP a.x, s OP a.y, s OP a.z, s OP a.w}; } \
    template <typename T, typename V> constexpr tvec2<T>& operator OP##=(tvec2<T>& a, const V& b) { a = a OP b; return a; }        \
    template <typename T, typename V> constexpr tvec3<T>& operator OP##=(tvec3<T>& a, const V& b) { a = a OP b; return a; }        \
    template <typename T, typename V> constexpr tvec4<T>& operator OP##=(tvec4<T>& a, const V& b) { a = a OP b; return a; }
CRT_GLM_BINOP(+)
CRT_GLM_BINOP(-)
CRT_GLM_BINOP(*)
CRT_GLM_BINOP(/)
#undef CRT_GLM_BINOP
// mixed scalar types (e.g. vec3 * double literal, vec3 * int): glm converts the scalar to T
#define CRT_GLM_MIXED(OP)                                                                                      \
    template <typename T, typename S, typename = std::enable_if_t<std::is_arithmetic_v<S> && !std::is_same_v<S, T>>> \
    constexpr tvec2<T> operator OP(const tvec2<T>& a, S s) { return a OP T(s); }                                     \
    template <typename T, typename S, typename = std::enable_if_t<std::is_arithmetic_v<S> && !std::is_same_v<S, T>>> \
    constexpr tvec2<T> operator OP(S s, const tvec2<T>& a) { return T(s) OP a; }                                     \
    template <typename T, typename S, typename = std::enable_if_t<std::is_arithmetic_v<S> && !std::is_same_v<S, T>>> \
    constexpr tvec3<T> operator OP(const tvec3<T>& a, S s) { return a OP T(s); }                                     \
    template <typename T, typename S, typename = std::enable_if_t<std::is_arithmetic_v<S> && !std::is_same_v<S, T>>> \
    constexpr tvec3<T> operator OP(S s, const tvec3<T>& a) { return T(s) OP a; }                                     \
    template <typename T, typename S, typename = std::enable_if_t<std::is_arithmetic_v<S> && !std::is_same_v<S, T>>> \
    constexpr tvec4<T> operator OP(const tvec4<T>& a, S s) { return a OP T(s); }                                     \
    template <typename T, typename S, typename = std::enable_if_t<std::is_arithmetic_v<S> && !std::is_same_v<S, T>>> \
    constexpr tvec4<T> operator OP(S s, const tvec4<T>& a) { return T(s) OP a; }
CRT_GLM_MIXED(+)
CRT_GLM_MIXED(-)
CRT_GLM_MIXED(*)
CRT_GLM_MIXED(/)
#undef CRT_GLM_MIXED

template <typename T> constexpr tvec2<T> operator-(const tvec2<T>& a) { return {-a.x, -a.y}; }
template <typename T> constexpr tvec3<T> operator-(const tvec3<T>& a) { return {-a.x, -a.y, -a.z}; }
template <typename T> constexpr tvec4<T> operator-(const tvec4<T>& a) { return {-a.x, -a.y, -a.z, -a.w}; }
template <typename T> constexpr bool operator==(const tvec2<T>& a, const tvec2<T>& b) { return a.x == b.x && a.y == b.y; }
template <typename T> constexpr bool operator!=(const tvec2<T>& a, const tvec2<T>& b) { return !(a == b); }
template <typename T> constexpr bool operator==(const tvec3<T>& a, const tvec3<T>& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
template <typename T> constexpr bool operator!=(const tvec3<T>& a, const tvec3<T>& b) { return !(a == b); }
template <typename T> constexpr bool operator==(const tvec4<T>& a, const tvec4<T>& b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }
template <typename T> constexpr bool operator!=(const tvec4<T>& a, const tvec4<T>& b) { return !(a == b); }

// ---- geometric functions (orders as in oracle/ovec.h) ---------------------------------------
template <typename T> constexpr T dot(const tvec2<T>& a, const tvec2<T>& b) { T px = a.x * b.x, py = a.y * b.y; return px + py; }
template <typename T> constexpr T dot(const tvec3<T>& a, const tvec3<T>& b) {
    T px = a.x * b.x, py = a.y * b.y, pz = a.z * b.z;
    return (px + py) + pz;
}
template <typename T> constexpr T dot(const tvec4<T>& a, const tvec4<T>& b) {
    T px = a.x * b.x, py = a.y * b.y, pz = a.z * b.z, pw = a.w * b.w;
    return (px + py) + (pz + pw);
}
template <typename T> constexpr tvec3<T> cross(const tvec3<T>& a, const tvec3<T>& b) {
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
template <typename T> inline T length(const tvec2<T>& v) { return std::sqrt(dot(v, v)); }
template <typename T> inline T length(const tvec3<T>& v) { return std::sqrt(dot(v, v)); }
template <typename T> inline T length(const tvec4<T>& v) { return std::sqrt(dot(v, v)); }
template <typename T> inline T distance(const tvec2<T>& a, const tvec2<T>& b) { return length(b - a); }
template <typename T> inline T distance(const tvec3<T>& a, const tvec3<T>& b) { return length(b - a); }
template <typename T> inline T inversesqrt(T x) { return T(1) / std::sqrt(x); }
template <typename T> inline tvec2<T> normalize(const tvec2<T>& v) { return v * inversesqrt(dot(v, v)); }
template <typename T> inline tvec3<T> normalize(const tvec3<T>& v) { return v * inversesqrt(dot(v, v)); }
template <typename T> inline tvec4<T> normalize(const tvec4<T>& v) { return v * inversesqrt(dot(v, v)); }

// ---- common functions ------------------------------------------------------------------------
template <typename T, typename = std::enable_if_t<std::is_arithmetic_v<T>>> constexpr T min(T x, T y) { return (y < x) ? y : x; }
template <typename T, typename = std::enable_if_t<std::is_arithmetic_v<T>>> constexpr T max(T x, T y) { return (x < y) ? y : x; }
template <typename T, typename = std::enable_if_t<std::is_arithmetic_v<T>>> constexpr T clamp(T x, T lo, T hi) { return min(max(x, lo), hi); }
template <typename T, typename = std::enable_if_t<std::is_arithmetic_v<T>>> constexpr T abs(T x) { return x < T(0) ? -x : x; }
template <typename T> constexpr tvec2<T> min(const tvec2<T>& a, const tvec2<T>& b) { return {min(a.x, b.x), min(a.y, b.y)}; }
template <typename T> constexpr tvec2<T> max(const tvec2<T>& a, const tvec2<T>& b) { return {max(a.x, b.x), max(a.y, b.y)}; }
template <typename T> constexpr tvec3<T> min(const tvec3<T>& a, const tvec3<T>& b) { return {min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)}; }
template <typename T> constexpr tvec3<T> max(const tvec3<T>& a, const tvec3<T>& b) { return {max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)}; }
template <typename T> constexpr tvec3<T> clamp(const tvec3<T>& v, T lo, T hi) { return {clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)}; }
template <typename T> constexpr tvec3<T> clamp(const tvec3<T>& v, const tvec3<T>& lo, const tvec3<T>& hi) { return {clamp(v.x, lo.x, hi.x), clamp(v.y, lo.y, hi.y), clamp(v.z, lo.z, hi.z)}; }
template <typename T> constexpr tvec2<T> clamp(const tvec2<T>& v, T lo, T hi) { return {clamp(v.x, lo, hi), clamp(v.y, lo, hi)}; }
template <typename T> constexpr tvec2<T> abs(const tvec2<T>& v) { return {abs(v.x), abs(v.y)}; }
template <typename T> constexpr tvec3<T> abs(const tvec3<T>& v) { return {abs(v.x), abs(v.y), abs(v.z)}; }
template <typename T> constexpr tvec4<T> abs(const tvec4<T>& v) { return {abs(v.x), abs(v.y), abs(v.z), abs(v.w)}; }
template <typename T> inline tvec2<T> floor(const tvec2<T>& v) { return {std::floor(v.x), std::floor(v.y)}; }
template <typename T> inline tvec3<T> floor(const tvec3<T>& v) { return {std::floor(v.x), std::floor(v.y), std::floor(v.z)}; }
template <typename T, typename U> constexpr T mix(T a, T b, U t) { return a * (U(1) - t) + b * t; }
template <typename T> constexpr T radians(T deg) { return deg * T(0.01745329251994329576923690768489); }
template <typename T> constexpr T degrees(T rad) { return rad * T(57.295779513082320876798154814105); }
template <typename T> constexpr T pi() { return T(3.14159265358979323846264338327950288); }
template <typename T> constexpr T epsilon() { return std::numeric_limits<T>::epsilon(); }
inline bool isnan(float x) { return std::isnan(x); }
inline bool isinf(float x) { return std::isinf(x); }

// ---- matrices (column-major: m[col][row]) ----------------------------------------------------
template <typename T> struct tmat4;
template <typename T> struct tmat3 {
    tvec3<T> c[3];
    constexpr tmat3() : c{tvec3<T>(1, 0, 0), tvec3<T>(0, 1, 0), tvec3<T>(0, 0, 1)} {}
    constexpr explicit tmat3(T s) : c{tvec3<T>(s, 0, 0), tvec3<T>(0, s, 0), tvec3<T>(0, 0, s)} {}
    constexpr tmat3(const tvec3<T>& a, const tvec3<T>& b, const tvec3<T>& d) : c{a, b, d} {}
    constexpr tmat3(T x0, T y0, T z0, T x1, T y1, T z1, T x2, T y2, T z2)
        : c{tvec3<T>(x0, y0, z0), tvec3<T>(x1, y1, z1), tvec3<T>(x2, y2, z2)} {}
    constexpr tmat3(const tmat4<T>& m);
    tvec3<T>& operator[](int i) { return c[i]; }
    constexpr const tvec3<T>& operator[](int i) const { return c[i]; }
};
template <typename T> struct tmat4 {
    tvec4<T> c[4];
    constexpr tmat4() : c{tvec4<T>(1, 0, 0, 0), tvec4<T>(0, 1, 0, 0), tvec4<T>(0, 0, 1, 0), tvec4<T>(0, 0, 0, 1)} {}
    constexpr explicit tmat4(T s) : c{tvec4<T>(s, 0, 0, 0), tvec4<T>(0, s, 0, 0), tvec4<T>(0, 0, s, 0), tvec4<T>(0, 0, 0, s)} {}
    constexpr tmat4(const tvec4<T>& a, const tvec4<T>& b, const tvec4<T>& d, const tvec4<T>& e) : c{a, b, d, e} {}
    constexpr tmat4(T x0, T y0, T z0, T w0, T x1, T y1, T z1, T w1, T x2, T y2, T z2, T w2, T x3, T y3, T z3, T w3)
        : c{tvec4<T>(x0, y0, z0, w0), tvec4<T>(x1, y1, z1, w1), tvec4<T>(x2, y2, z2, w2), tvec4<T>(x3, y3, z3, w3)} {}
    constexpr explicit tmat4(const tmat3<T>& m)
        : c{tvec4<T>(m[0], 0), tvec4<T>(m[1], 0), tvec4<T>(m[2], 0), tvec4<T>(0, 0, 0, 1)} {}
    tvec4<T>& operator[](int i) { return c[i]; }
    constexpr const tvec4<T>& operator[](int i) const { return c[i]; }
};
template <typename T> constexpr tmat3<T>::tmat3(const tmat4<T>& m) : c{tvec3<T>(m[0]), tvec3<T>(m[1]), tvec3<T>(m[2])} {}
typedef tmat3<float> mat3; typedef tmat4<float> mat4; typedef tmat3<double> dmat3; typedef tmat4<double> dmat4;
typedef mat3 mat3x3; typedef mat4 mat4x4;

// mat4 * vec4, glm scalar path: (m0*x + m1*y) + (m2*z + m3*w)
template <typename T> constexpr tvec4<T> operator*(const tmat4<T>& m, const tvec4<T>& v) {
    tvec4<T> a = m[0] * v.x + m[1] * v.y;
    tvec4<T> b = m[2] * v.z + m[3] * v.w;
    return a + b;
}
// mat3 * vec3: per row, left to right
template <typename T> constexpr tvec3<T> operator*(const tmat3<T>& m, const tvec3<T>& v) {
    return {(m[0][0] * v.x + m[1][0] * v.y) + m[2][0] * v.z,
            (m[0][1] * v.x + m[1][1] * v.y) + m[2][1] * v.z,
            (m[0][2] * v.x + m[1][2] * v.y) + m[2][2] * v.z};
}
template <typename T> constexpr tmat4<T> operator*(const tmat4<T>& a, const tmat4<T>& b) {
    tmat4<T> r;
    for (int j = 0; j < 4; ++j) r[j] = ((a[0] * b[j].x + a[1] * b[j].y) + a[2] * b[j].z) + a[3] * b[j].w;
    return r;
}
template <typename T> constexpr tmat3<T> operator*(const tmat3<T>& a, const tmat3<T>& b) {
    tmat3<T> r;
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) r[j][i] = (a[0][i] * b[j][0] + a[1][i] * b[j][1]) + a[2][i] * b[j][2];
    return r;
}
template <typename T> constexpr tmat4<T> operator*(const tmat4<T>& a, T s) { return {a[0] * s, a[1] * s, a[2] * s, a[3] * s}; }
template <typename T> constexpr tmat3<T> transpose(const tmat3<T>& m) {
    tmat3<T> r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r[i][j] = m[j][i];
    return r;
}
template <typename T> constexpr tmat4<T> transpose(const tmat4<T>& m) {
    tmat4<T> r;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r[i][j] = m[j][i];
    return r;
}
template <typename T> constexpr tmat3<T> inverse(const tmat3<T>& m) {
    T d = T(1) / (+m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2])
                  - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2])
                  + m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]));
    tmat3<T> r;
    r[0][0] = +(m[1][1] * m[2][2] - m[2][1] * m[1][2]) * d;
    r[1][0] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]) * d;
    r[2][0] = +(m[1][0] * m[2][1] - m[2][0] * m[1][1]) * d;
    r[0][1] = -(m[0][1] * m[2][2] - m[2][1] * m[0][2]) * d;
    r[1][1] = +(m[0][0] * m[2][2] - m[2][0] * m[0][2]) * d;
    r[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]) * d;
    r[0][2] = +(m[0][1] * m[1][2] - m[1][1] * m[0][2]) * d;
    r[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]) * d;
    r[2][2] = +(m[0][0] * m[1][1] - m[1][0] * m[0][1]) * d;
    return r;
}
// cofactor inverse, sub-factors grouped as glm's compute_inverse<4,4> does
template <typename T> constexpr tmat4<T> inverse(const tmat4<T>& m) {
    T c00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
    T c02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
    T c03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
    T c04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
    T c06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
    T c07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
    T c08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
    T c10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
    T c11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
    T c12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
    T c14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
    T c15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
    T c16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
    T c18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
    T c19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
    T c20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    T c22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
    T c23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
    tvec4<T> f0(c00, c00, c02, c03), f1(c04, c04, c06, c07), f2(c08, c08, c10, c11);
    tvec4<T> f3(c12, c12, c14, c15), f4(c16, c16, c18, c19), f5(c20, c20, c22, c23);
    tvec4<T> v0(m[1][0], m[0][0], m[0][0], m[0][0]), v1(m[1][1], m[0][1], m[0][1], m[0][1]);
    tvec4<T> v2(m[1][2], m[0][2], m[0][2], m[0][2]), v3(m[1][3], m[0][3], m[0][3], m[0][3]);
    tvec4<T> i0 = (v1 * f0 - v2 * f1) + v3 * f2;
    tvec4<T> i1 = (v0 * f0 - v2 * f3) + v3 * f4;
    tvec4<T> i2 = (v0 * f1 - v1 * f3) + v3 * f5;
    tvec4<T> i3 = (v0 * f2 - v1 * f4) + v2 * f5;
    tvec4<T> sa(+1, -1, +1, -1), sb(-1, +1, -1, +1);
    tmat4<T> inv(i0 * sa, i1 * sb, i2 * sa, i3 * sb);
    tvec4<T> row0(inv[0][0], inv[1][0], inv[2][0], inv[3][0]);
    tvec4<T> d0 = m[0] * row0;
    T det = (d0.x + d0.y) + (d0.z + d0.w);
    T ood = T(1) / det;
    return inv * ood;
}
template <typename T> constexpr tmat4<T> translate(const tmat4<T>& m, const tvec3<T>& v) {
    tmat4<T> r = m;
    r[3] = ((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3];
    return r;
}
template <typename T> constexpr tmat4<T> scale(const tmat4<T>& m, const tvec3<T>& v) {
    return {m[0] * v.x, m[1] * v.y, m[2] * v.z, m[3]};
}
template <typename T> inline tmat4<T> rotate(const tmat4<T>& m, T angle, const tvec3<T>& v) {
    T c = std::cos(angle), s = std::sin(angle);
    tvec3<T> axis = normalize(v);
    tvec3<T> t = axis * (T(1) - c);
    T R[3][3];
    R[0][0] = c + t.x * axis.x;          R[0][1] = t.x * axis.y + s * axis.z; R[0][2] = t.x * axis.z - s * axis.y;
    R[1][0] = t.y * axis.x - s * axis.z; R[1][1] = c + t.y * axis.y;          R[1][2] = t.y * axis.z + s * axis.x;
    R[2][0] = t.z * axis.x + s * axis.y; R[2][1] = t.z * axis.y - s * axis.x; R[2][2] = c + t.z * axis.z;
    tmat4<T> r;
    for (int j = 0; j < 3; ++j) r[j] = (m[0] * R[j][0] + m[1] * R[j][1]) + m[2] * R[j][2];
    r[3] = m[3];
    return r;
}
template <typename T> constexpr tmat4<T> diagonal4x4(const tvec4<T>& v) {
    return {tvec4<T>(v.x, 0, 0, 0), tvec4<T>(0, v.y, 0, 0), tvec4<T>(0, 0, v.z, 0), tvec4<T>(0, 0, 0, v.w)};
}
template <typename T> constexpr tmat3<T> diagonal3x3(const tvec3<T>& v) {
    return {tvec3<T>(v.x, 0, 0), tvec3<T>(0, v.y, 0), tvec3<T>(0, 0, v.z)};
}

// ---- pointers / strings ------------------------------------------------------------------------
template <typename T> inline T* value_ptr(tvec2<T>& v) { return &v.x; }
template <typename T> inline T* value_ptr(tvec3<T>& v) { return &v.x; }
template <typename T> inline T* value_ptr(tvec4<T>& v) { return &v.x; }
template <typename T> inline T* value_ptr(tmat3<T>& m) { return &m[0].x; }
template <typename T> inline T* value_ptr(tmat4<T>& m) { return &m[0].x; }
template <typename T> inline const T* value_ptr(const tmat4<T>& m) { return &m[0].x; }
template <typename T> inline std::string to_string(const tvec2<T>& v) { return "vec2(" + std::to_string(v.x) + ", " + std::to_string(v.y) + ")"; }
template <typename T> inline std::string to_string(const tvec3<T>& v) { return "vec3(" + std::to_string(v.x) + ", " + std::to_string(v.y) + ", " + std::to_string(v.z) + ")"; }
template <typename T> inline std::string to_string(const tvec4<T>& v) { return "vec4(" + std::to_string(v.x) + ", " + std::to_string(v.y) + ", " + std::to_string(v.z) + ", " + std::to_string(v.w) + ")"; }
template <typename T> inline std::string to_string(const tmat3<T>& m) { return "mat3(" + to_string(m[0]) + ", " + to_string(m[1]) + ", " + to_string(m[2]) + ")"; }
template <typename T> inline std::string to_string(const tmat4<T>& m) { return "mat4(" + to_string(m[0]) + ", " + to_string(m[1]) + ", " + to_string(m[2]) + ", " + to_string(m[3]) + ")"; }

}  // namespace glm
