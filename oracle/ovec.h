// ORACLE (test infrastructure, never shipped, never on the product path).
//
// Minimal fp32 vector/matrix layer standing in for glm, which the reference uses
// everywhere (pch.h:25-30) but does not vendor or pin.  Every function fixes ONE
// evaluation order (the one glm's scalar code path uses) so the oracle, built with
// -ffp-contract=off, is reproducible.  Nothing in the reference pins results at this
// boundary (glm is neither vendored nor versioned): this is the one place
// where parity stays "unpinned".  oracle/refshim/glm/glm.hpp, against which the reference's
// own sources are compiled (oracle/_ref), uses the same orders; matrices are computed once
// on the host and handed to both oracle and GPU, so only the hot-path orders below matter.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace orc {

struct vec2 {
    float x = 0, y = 0;
    vec2() = default;
    vec2(float a, float b) : x(a), y(b) {}
    float operator[](int i) const { return i == 0 ? x : y; }
};
struct ivec2 {
    int x = 0, y = 0;
    ivec2() = default;
    ivec2(int a, int b) : x(a), y(b) {}
};
struct vec3 {
    float x = 0, y = 0, z = 0;
    vec3() = default;
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
struct dvec3 {
    double x = 0, y = 0, z = 0;
};
struct vec4 {
    float x = 0, y = 0, z = 0, w = 0;
    vec4() = default;
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(vec3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
};

inline vec2 operator+(vec2 a, vec2 b) { return {a.x + b.x, a.y + b.y}; }
inline vec2 operator-(vec2 a, vec2 b) { return {a.x - b.x, a.y - b.y}; }
inline vec2 operator*(vec2 a, vec2 b) { return {a.x * b.x, a.y * b.y}; }
inline vec2 operator*(float s, vec2 a) { return {s * a.x, s * a.y}; }
inline vec2 operator*(vec2 a, float s) { return {a.x * s, a.y * s}; }
inline vec2 operator/(vec2 a, vec2 b) { return {a.x / b.x, a.y / b.y}; }

inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline vec3 operator/(vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }

inline vec4 operator+(vec4 a, vec4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline vec4 operator*(vec4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline vec4 operator*(vec4 a, vec4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }

inline vec3 xyz(vec4 v) { return {v.x, v.y, v.z}; }

// glm::dot: products first, then left-to-right sum (vec3) / pairwise (vec4)
inline float dot(vec2 a, vec2 b) { float px = a.x * b.x, py = a.y * b.y; return px + py; }
inline float dot(vec3 a, vec3 b) {
    float px = a.x * b.x, py = a.y * b.y, pz = a.z * b.z;
    return (px + py) + pz;
}
inline float dot(vec4 a, vec4 b) {
    float px = a.x * b.x, py = a.y * b.y, pz = a.z * b.z, pw = a.w * b.w;
    return (px + py) + (pz + pw);
}
// glm::cross(x, y)
inline vec3 cross(vec3 a, vec3 b) {
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
inline dvec3 cross(dvec3 a, dvec3 b) {
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
inline float length(vec3 v) { return std::sqrt(dot(v, v)); }
inline float length(vec2 v) { return std::sqrt(dot(v, v)); }
inline float distance(vec3 a, vec3 b) { return length(b - a); }
// glm::normalize = v * inversesqrt(dot(v,v)), inversesqrt(x) = 1/sqrt(x)
inline vec3 normalize(vec3 v) { float s = 1.0f / std::sqrt(dot(v, v)); return v * s; }
inline vec4 normalize(vec4 v) { float s = 1.0f / std::sqrt(dot(v, v)); return v * s; }
// glm::min/max/clamp (NaN behaviour of the ternaries preserved)
inline float gmin(float x, float y) { return (y < x) ? y : x; }
inline float gmax(float x, float y) { return (x < y) ? y : x; }
inline float clampf(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }

// column-major like glm: m.c[col][row]
struct mat3 {
    float c[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    vec3 col(int i) const { return {c[i][0], c[i][1], c[i][2]}; }
    void setcol(int i, vec3 v) { c[i][0] = v.x; c[i][1] = v.y; c[i][2] = v.z; }
};
struct mat4 {
    float c[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    vec4 col(int i) const { return {c[i][0], c[i][1], c[i][2], c[i][3]}; }
    void setcol(int i, vec4 v) { c[i][0] = v.x; c[i][1] = v.y; c[i][2] = v.z; c[i][3] = v.w; }
    static mat4 from_cols(vec4 a, vec4 b, vec4 d, vec4 e) {
        mat4 m; m.setcol(0, a); m.setcol(1, b); m.setcol(2, d); m.setcol(3, e); return m;
    }
    static mat4 from_ptr(const float* p) { mat4 m; std::memcpy(m.c, p, 64); return m; }
};

// glm mat4 * vec4 (scalar path): (m0*x + m1*y) + (m2*z + m3*w)
inline vec4 mul(const mat4& m, vec4 v) {
    vec4 a = m.col(0) * v.x + m.col(1) * v.y;
    vec4 b = m.col(2) * v.z + m.col(3) * v.w;
    return a + b;
}
// glm mat3 * vec3: m00*x + m10*y + m20*z, left to right
inline vec3 mul(const mat3& m, vec3 v) {
    return {(m.c[0][0] * v.x + m.c[1][0] * v.y) + m.c[2][0] * v.z,
            (m.c[0][1] * v.x + m.c[1][1] * v.y) + m.c[2][1] * v.z,
            (m.c[0][2] * v.x + m.c[1][2] * v.y) + m.c[2][2] * v.z};
}
inline mat4 mul(const mat4& a, const mat4& b) {
    mat4 r;
    for (int j = 0; j < 4; ++j) {
        vec4 bj = b.col(j);
        vec4 t = ((a.col(0) * bj.x + a.col(1) * bj.y) + a.col(2) * bj.z) + a.col(3) * bj.w;
        r.setcol(j, t);
    }
    return r;
}
inline mat3 mul(const mat3& a, const mat3& b) {
    mat3 r;
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i)
            r.c[j][i] = (a.c[0][i] * b.c[j][0] + a.c[1][i] * b.c[j][1]) + a.c[2][i] * b.c[j][2];
    return r;
}
inline mat3 transpose(const mat3& m) {
    mat3 r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.c[i][j] = m.c[j][i];
    return r;
}
inline mat3 upper3(const mat4& m) {
    mat3 r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.c[i][j] = m.c[i][j];
    return r;
}
inline mat3 inverse(const mat3& m) {
    float d = 1.0f / (+m.c[0][0] * (m.c[1][1] * m.c[2][2] - m.c[2][1] * m.c[1][2])
                      - m.c[1][0] * (m.c[0][1] * m.c[2][2] - m.c[2][1] * m.c[0][2])
                      + m.c[2][0] * (m.c[0][1] * m.c[1][2] - m.c[1][1] * m.c[0][2]));
    mat3 r;
    r.c[0][0] = +(m.c[1][1] * m.c[2][2] - m.c[2][1] * m.c[1][2]) * d;
    r.c[1][0] = -(m.c[1][0] * m.c[2][2] - m.c[2][0] * m.c[1][2]) * d;
    r.c[2][0] = +(m.c[1][0] * m.c[2][1] - m.c[2][0] * m.c[1][1]) * d;
    r.c[0][1] = -(m.c[0][1] * m.c[2][2] - m.c[2][1] * m.c[0][2]) * d;
    r.c[1][1] = +(m.c[0][0] * m.c[2][2] - m.c[2][0] * m.c[0][2]) * d;
    r.c[2][1] = -(m.c[0][0] * m.c[2][1] - m.c[2][0] * m.c[0][1]) * d;
    r.c[0][2] = +(m.c[0][1] * m.c[1][2] - m.c[1][1] * m.c[0][2]) * d;
    r.c[1][2] = -(m.c[0][0] * m.c[1][2] - m.c[1][0] * m.c[0][2]) * d;
    r.c[2][2] = +(m.c[0][0] * m.c[1][1] - m.c[1][0] * m.c[0][1]) * d;
    return r;
}
// cofactor inverse in glm's sub-factor order
inline mat4 inverse(const mat4& m) {
    auto M = [&](int col, int row) { return m.c[col][row]; };
    float c00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3);
    float c02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3);
    float c03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
    float c04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
    float c06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3);
    float c07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
    float c08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2);
    float c10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2);
    float c11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
    float c12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
    float c14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3);
    float c15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
    float c16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2);
    float c18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2);
    float c19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
    float c20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
    float c22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1);
    float c23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
    vec4 f0(c00, c00, c02, c03), f1(c04, c04, c06, c07), f2(c08, c08, c10, c11);
    vec4 f3(c12, c12, c14, c15), f4(c16, c16, c18, c19), f5(c20, c20, c22, c23);
    vec4 v0(M(1, 0), M(0, 0), M(0, 0), M(0, 0)), v1(M(1, 1), M(0, 1), M(0, 1), M(0, 1));
    vec4 v2(M(1, 2), M(0, 2), M(0, 2), M(0, 2)), v3(M(1, 3), M(0, 3), M(0, 3), M(0, 3));
    auto sub = [](vec4 a, vec4 b) { return vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); };
    vec4 i0 = sub(v1 * f0, v2 * f1) + v3 * f2;
    vec4 i1 = sub(v0 * f0, v2 * f3) + v3 * f4;
    vec4 i2 = sub(v0 * f1, v1 * f3) + v3 * f5;
    vec4 i3 = sub(v0 * f2, v1 * f4) + v2 * f5;
    vec4 sa(+1, -1, +1, -1), sb(-1, +1, -1, +1);
    mat4 inv = mat4::from_cols(i0 * sa, i1 * sb, i2 * sa, i3 * sb);
    vec4 row0(inv.c[0][0], inv.c[1][0], inv.c[2][0], inv.c[3][0]);
    vec4 d0 = m.col(0) * row0;
    float det = (d0.x + d0.y) + (d0.z + d0.w);
    float ood = 1.0f / det;
    mat4 r;
    for (int j = 0; j < 4; ++j) r.setcol(j, inv.col(j) * ood);
    return r;
}
inline mat4 translate(const mat4& m, vec3 v) {
    mat4 r = m;
    r.setcol(3, ((m.col(0) * v.x + m.col(1) * v.y) + m.col(2) * v.z) + m.col(3));
    return r;
}
inline mat4 scale(const mat4& m, vec3 v) {
    mat4 r;
    r.setcol(0, m.col(0) * v.x); r.setcol(1, m.col(1) * v.y); r.setcol(2, m.col(2) * v.z); r.setcol(3, m.col(3));
    return r;
}
inline mat4 rotate(const mat4& m, float angle, vec3 v) {
    float c = std::cos(angle), s = std::sin(angle);
    vec3 axis = normalize(v);
    vec3 t = axis * (1.0f - c);
    float R[3][3];
    R[0][0] = c + t.x * axis.x;          R[0][1] = t.x * axis.y + s * axis.z; R[0][2] = t.x * axis.z - s * axis.y;
    R[1][0] = t.y * axis.x - s * axis.z; R[1][1] = c + t.y * axis.y;          R[1][2] = t.y * axis.z + s * axis.x;
    R[2][0] = t.z * axis.x + s * axis.y; R[2][1] = t.z * axis.y - s * axis.x; R[2][2] = c + t.z * axis.z;
    mat4 r;
    for (int j = 0; j < 3; ++j)
        r.setcol(j, (m.col(0) * R[j][0] + m.col(1) * R[j][1]) + m.col(2) * R[j][2]);
    r.setcol(3, m.col(3));
    return r;
}

}  // namespace orc
