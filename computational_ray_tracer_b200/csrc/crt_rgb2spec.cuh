// crt_rgb2spec.cuh -- the sRGB -> sigmoid-polynomial coefficient table behind RGBToSpectrumTable (ThirdParty/pbrv4/
// color.h:405-432, color.cpp:26-73).
//
// The reference READS this table from `../rgb2spec/sRGB64binary` (color.cpp:107-166), a file that is not in its
// repository, so until now only grey RGB (which bypasses the table, color.cpp:35-37) could be rendered.
// This header regenerates the table.  The generating tool is third-party (pbrt-v4's `rgb2spec_opt`, Jakob & Hanika 2019,
// "A Low-Dimensional Function Space for Efficient Spectral Upsampling"; not vendored, version unpinned), so its published
// algorithm is restated here from scratch:
//   for every table cell (max component l, brightness scale[k], chromaticities x = i/63, y = j/63) fit (A, B, C) such that
//   s(A t^2 + B t + C), t = (lambda-360)/470, seen under the colour space's illuminant, has the cell's RGB; Gauss-Newton on
//   the CIELAB residual, <= 15 iterations, central-difference Jacobian (eps 1e-4), coefficients clamped to |c| <= 200,
//   warm-started along k from k = res/5 upwards and downwards; the result is re-expressed for lambda in nm.
// Differences from the tool, both deliberate: the spectral integrals are plain sums over the 471 integer wavelengths -- the
// quadrature the renderer itself uses for every spectrum -> XYZ conversion (InnerProduct, spectrum.h:762-768) -- instead of
// Simpson 3/8 on 283 interpolated points, and the CIE / D65 data are the reference's own tables (spectrum.cpp:298-2600).
// The LOOKUP (trilinear interpolation, rgb2spec_lookup below) follows color.cpp:26-73 operation by operation and is pinned
// bit for bit against the reference's compiled code (tests/test_cpu_rgb2spec.py); the generator is validated by round trips.
//
// One warp fits one chain of 64 cells: the lanes split the 471 wavelengths, partial sums are combined with XOR shuffles so
// every lane holds identical sums and the (tiny) Gauss-Newton step is computed redundantly -- no divergence, no shared
// memory.  3 x 64 x 64 = 12 288 independent warps; fp64 throughout (B200 has full-rate-enough fp64; the whole table takes
// a few tens of ms).  The same code runs on the host (one "lane") for crt_rgb2spec_fit, used by the CPU tests.
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define R2S_HD __host__ __device__
#else
#define R2S_HD
#endif

namespace crt {

constexpr int kRgb2SpecRes = 64;              // RGBToSpectrumTable::res, color.h:410 (the reference's lookup hard-codes 64)
constexpr int kRgb2SpecSamples = 471;         // 360..830 nm

struct Rgb2SpecModel {
    double rgb_tbl[3][kRgb2SpecSamples];      // RGBFromXYZ * (xbar,ybar,zbar)(lambda) * illuminant(lambda)
    double white[3];                          // XYZ of the illuminant under the same quadrature
    double rgb_to_xyz[9];                     // column-major like glm: m[3*col+row]
};

R2S_HD inline double r2s_smoothstep(double x) { return x * x * (3.0 - 2.0 * x); }
R2S_HD inline double r2s_scale(int k) { return r2s_smoothstep(r2s_smoothstep((double)k / (kRgb2SpecRes - 1))); }

R2S_HD inline void r2s_cie_lab(const Rgb2SpecModel& M, double* p) {
    double X = 0, Y = 0, Z = 0;
    for (int j = 0; j < 3; ++j) {
        X += p[j] * M.rgb_to_xyz[3 * j + 0];
        Y += p[j] * M.rgb_to_xyz[3 * j + 1];
        Z += p[j] * M.rgb_to_xyz[3 * j + 2];
    }
    auto f = [](double t) {
        const double delta = 6.0 / 29.0;
        return t > delta * delta * delta ? cbrt(t) : t / (delta * delta * 3.0) + (4.0 / 29.0);
    };
    double fx = f(X / M.white[0]), fy = f(Y / M.white[1]), fz = f(Z / M.white[2]);
    p[0] = 116.0 * fy - 16.0;
    p[1] = 500.0 * (fx - fy);
    p[2] = 200.0 * (fy - fz);
}

// RGB of the sigmoid spectrum with normalised-wavelength coefficients c, mapped to CIELAB.  LANES = 32: called by a full
// warp, lane = threadIdx.x & 31; LANES = 1: sequential.
template <int LANES>
R2S_HD inline void r2s_eval_lab(const Rgb2SpecModel& M, const double* c, int lane, double* lab) {
    double out[3] = {0, 0, 0};
    for (int i = lane; i < kRgb2SpecSamples; i += LANES) {
        double t = (double)i / (kRgb2SpecSamples - 1);
        double x = (c[0] * t + c[1]) * t + c[2];
        double y = 1.0 / sqrt(x * x + 1.0);
        double s = 0.5 * x * y + 0.5;
        out[0] += M.rgb_tbl[0][i] * s; out[1] += M.rgb_tbl[1][i] * s; out[2] += M.rgb_tbl[2][i] * s;
    }
#ifdef __CUDA_ARCH__
    if (LANES == 32) {
        for (int o = 16; o > 0; o >>= 1)
            for (int j = 0; j < 3; ++j) out[j] += __shfl_xor_sync(0xffffffffu, out[j], o);
    }
#endif
    r2s_cie_lab(M, out);
    lab[0] = out[0]; lab[1] = out[1]; lab[2] = out[2];
}

// 3x3 solve by LU with partial pivoting; false if singular
R2S_HD inline bool r2s_solve3(double A[3][3], const double* b, double* x) {
    int P[3] = {0, 1, 2};
    for (int i = 0; i < 3; ++i) {
        double maxA = 0; int imax = i;
        for (int k = i; k < 3; ++k) { double a = fabs(A[k][i]); if (a > maxA) { maxA = a; imax = k; } }
        if (maxA < 1e-15) return false;
        if (imax != i) {
            int t = P[i]; P[i] = P[imax]; P[imax] = t;
            for (int k = 0; k < 3; ++k) { double tt = A[i][k]; A[i][k] = A[imax][k]; A[imax][k] = tt; }
        }
        for (int j = i + 1; j < 3; ++j) {
            A[j][i] /= A[i][i];
            for (int k = i + 1; k < 3; ++k) A[j][k] -= A[j][i] * A[i][k];
        }
    }
    for (int i = 0; i < 3; ++i) { x[i] = b[P[i]]; for (int k = 0; k < i; ++k) x[i] -= A[i][k] * x[k]; }
    for (int i = 2; i >= 0; --i) { for (int k = i + 1; k < 3; ++k) x[i] -= A[i][k] * x[k]; x[i] /= A[i][i]; }
    return true;
}

template <int LANES>
R2S_HD inline void r2s_gauss_newton(const Rgb2SpecModel& M, const double* rgb, double* c, int lane) {
    double target[3] = {rgb[0], rgb[1], rgb[2]};
    r2s_cie_lab(M, target);
    for (int it = 0; it < 15; ++it) {
        double lab[3], res[3], J[3][3];
        r2s_eval_lab<LANES>(M, c, lane, lab);
        for (int j = 0; j < 3; ++j) res[j] = target[j] - lab[j];
        const double eps = 1e-4;
        for (int i = 0; i < 3; ++i) {
            double tmp[3] = {c[0], c[1], c[2]}, r0[3], r1[3];
            tmp[i] = c[i] - eps; r2s_eval_lab<LANES>(M, tmp, lane, r0);
            tmp[i] = c[i] + eps; r2s_eval_lab<LANES>(M, tmp, lane, r1);
            // d(residual)/dc = -d(lab)/dc
            for (int j = 0; j < 3; ++j) J[j][i] = -(r1[j] - r0[j]) * (1.0 / (2 * eps));
        }
        double x[3];
        if (!r2s_solve3(J, res, x)) return;
        double r = 0;
        for (int j = 0; j < 3; ++j) { c[j] -= x[j]; r += res[j] * res[j]; }
        double mx = fmax(fmax(fabs(c[0]), fabs(c[1])), fabs(c[2]));
        if (mx > 200.0) { double s = 200.0 / mx; c[0] *= s; c[1] *= s; c[2] *= s; }
        if (r < 1e-6) break;
    }
}

// normalised-wavelength coefficients -> coefficients of lambda in nm, as stored in the table
R2S_HD inline void r2s_to_nm(const double* c, float* out) {
    const double c0 = 360.0, c1 = 1.0 / (830.0 - 360.0);
    double A = c[0], B = c[1], C = c[2];
    out[0] = (float)(A * (c1 * c1));
    out[1] = (float)(B * c1 - 2 * A * c0 * (c1 * c1));
    out[2] = (float)(C - B * c0 * c1 + A * (c0 * c1) * (c0 * c1));
}

// One chain = all 64 brightness levels of (l, j, i).  data layout: [l][k][j][i][3] (color.cpp:57: maxc, z, y, x, coefficient)
template <int LANES>
R2S_HD inline void r2s_chain(const Rgb2SpecModel& M, int l, int j, int i, int lane, float* data) {
    const int res = kRgb2SpecRes;
    const double x = (double)i / (res - 1), y = (double)j / (res - 1);
    const int start = res / 5;
    for (int pass = 0; pass < 2; ++pass) {
        double c[3] = {0, 0, 0};
        for (int k = start; pass == 0 ? k < res : k >= 0; k += pass == 0 ? 1 : -1) {
            double b = r2s_scale(k), rgb[3];
            rgb[l] = b; rgb[(l + 1) % 3] = x * b; rgb[(l + 2) % 3] = y * b;
            r2s_gauss_newton<LANES>(M, rgb, c, lane);
            if (lane == 0) r2s_to_nm(c, data + 3 * ((((size_t)l * res + k) * res + j) * res + i));
        }
    }
}

// RGBToSpectrumTable::operator() (color.cpp:26-73), operation by operation.  zNodes = scale (64), coeffs = data.
// Host only: spectra are built at scene-setup time; the device evaluates the resulting sigmoid polynomial.
inline void rgb2spec_lookup(const float* zNodes, const float* coeffs, const float* rgb, float* c_out) {
    const int res = kRgb2SpecRes;
    if (rgb[0] == rgb[1] && rgb[1] == rgb[2]) {                        // color.cpp:35-37
        c_out[0] = 0; c_out[1] = 0; c_out[2] = (rgb[0] - .5f) / std::sqrt(rgb[0] * (1 - rgb[0]));
        return;
    }
    int maxc = (rgb[0] > rgb[1]) ? ((rgb[0] > rgb[2]) ? 0 : 2) : ((rgb[1] > rgb[2]) ? 1 : 2);
    float z = rgb[maxc];
    float x = rgb[(maxc + 1) % 3] * (res - 1) / z;
    float y = rgb[(maxc + 2) % 3] * (res - 1) / z;
    int xi = std::min((int)x, res - 2), yi = std::min((int)y, res - 2);
    // FindInterval(res, zNodes[i] < z), helpers.h:160-172
    long size = (long)res - 2, first = 1;
    while (size > 0) {
        long half = size >> 1, middle = first + half;
        bool pr = zNodes[middle] < z;
        first = pr ? middle + 1 : first;
        size = pr ? size - (half + 1) : half;
    }
    long v = first - 1, hi = (long)res - 2;
    int zi = (int)(v < 0 ? 0 : (v > hi ? hi : v));
    float dx = x - xi, dy = y - yi, dz = (z - zNodes[zi]) / (zNodes[zi + 1] - zNodes[zi]);
    auto lerp = [](float t, float a, float b) { return (1 - t) * a + t * b; };       // helpers.h:154-157
    for (int i = 0; i < 3; ++i) {
        auto co = [&](int ddx, int ddy, int ddz) {
            return coeffs[(size_t)maxc * 64 * 64 * 64 * 3 + (size_t)(zi + ddz) * 64 * 64 * 3 + (size_t)(yi + ddy) * 64 * 3 + (size_t)(xi + ddx) * 3 + i];
        };
        c_out[i] = lerp(dz, lerp(dy, lerp(dx, co(0, 0, 0), co(1, 0, 0)), lerp(dx, co(0, 1, 0), co(1, 1, 0))),
                        lerp(dy, lerp(dx, co(0, 0, 1), co(1, 0, 1)), lerp(dx, co(0, 1, 1), co(1, 1, 1))));
    }
}

}  // namespace crt
